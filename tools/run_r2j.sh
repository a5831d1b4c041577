mkdir -p gpurun_out/r2j
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2j/pytest_gpu.txt; cat gpurun_out/r2j/pytest_gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29601 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2j/bench_a_n2.json 2> gpurun_out/r2j/bench_a_n2.err; tail -c 1500 gpurun_out/r2j/bench_a_n2.json; tail -3 gpurun_out/r2j/bench_a_n2.err
python bench.py --impl reference --steps 1 > gpurun_out/r2j/bench_reference.json 2> gpurun_out/r2j/bench_reference.err; cat gpurun_out/r2j/bench_reference.json | head -c 1200
$TR --master-port 29602 tools/par_check.py 6 4 4 3 > gpurun_out/r2j/par_check_p3_n2.txt 2>&1; tail -2 gpurun_out/r2j/par_check_p3_n2.txt
SB_BENCH_CONFIG=b SB_BENCH_B_N=24 $TR --master-port 29603 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2j/bench_b_small_n2.json 2> gpurun_out/r2j/bench_b_small_n2.err; tail -c 800 gpurun_out/r2j/bench_b_small_n2.json; tail -3 gpurun_out/r2j/bench_b_small_n2.err
SB_BENCH_CONFIG=b $TR --master-port 29604 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2j/bench_b_weak_n2.json 2> gpurun_out/r2j/bench_b_weak_n2.err; tail -c 600 gpurun_out/r2j/bench_b_weak_n2.json; tail -3 gpurun_out/r2j/bench_b_weak_n2.err
SB_BENCH_CONFIG=b SB_BENCH_SCALING=strong $TR --master-port 29605 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2j/bench_b_strong_n2.json 2> gpurun_out/r2j/bench_b_strong_n2.err; tail -c 600 gpurun_out/r2j/bench_b_strong_n2.json; tail -3 gpurun_out/r2j/bench_b_strong_n2.err
