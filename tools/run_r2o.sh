mkdir -p gpurun_out/r2o
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_halo.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2 3 4; do
  SB_BENCH_CONFIG=b SB_BENCH_B_N=48 $TR --master-port 2962$i bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2o/b48_$i.json 2> gpurun_out/r2o/b48_$i.err; grep -h "AssertionError: rank" gpurun_out/r2o/b48_$i.err | head -2; tail -c 200 gpurun_out/r2o/b48_$i.json | head -c 150; echo
done
for i in 1 2; do
  SB_BENCH_CONFIG=b SB_BENCH_SCALING=strong $TR --master-port 2963$i bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2o/bstrong_$i.json 2> gpurun_out/r2o/bstrong_$i.err; grep -h "AssertionError: rank" gpurun_out/r2o/bstrong_$i.err | head -2; tail -c 200 gpurun_out/r2o/bstrong_$i.json | head -c 150; echo
done
