mkdir -p gpurun_out/r2m
python -m pytest tests/test_gpu_cpp_multirank.py tests/test_gpu_halo.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2m/pytest_multirank.txt; cat gpurun_out/r2m/pytest_multirank.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29611 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r2m/bench_a_n2.json 2> gpurun_out/r2m/bench_a_n2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2m/bench_a_n2.json').read().strip().splitlines()[-1])
print('N=2 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'], 'parity', d.get('parity_vs_global'))
PY
tail -3 gpurun_out/r2m/bench_a_n2.err
python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2m/bench_a_n1.json 2> gpurun_out/r2m/bench_a_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2m/bench_a_n1.json').read().strip().splitlines()[-1])
print('N=1 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'])
PY
