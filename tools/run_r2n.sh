mkdir -p gpurun_out/r2n
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2n/pytest_gpu.txt; cat gpurun_out/r2n/pytest_gpu.txt
for ell in 1 0; do
  SB_SCATTER_ELL=$ell python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-solver-layer > gpurun_out/r2n/bench_ell$ell.json 2>gpurun_out/r2n/bench_ell$ell.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2n/bench_ell$ell.json').read().strip().splitlines()[-1])
print('ELL=$ell', 'value', round(d['value'],4), 'ms', round(d['ms_per_step'],5), d['phases_ms'])
PY
  SB_SCATTER_ELL=$ell python tools/config_b.py 96 > gpurun_out/r2n/config_b96_ell$ell.json 2>&1; grep -E "scatter|step_ms|PASS|FAIL" gpurun_out/r2n/config_b96_ell$ell.json
done
