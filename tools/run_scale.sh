# usage: bash tools/run_scale.sh N tag     (N GPUs of this box; results under gpurun_out/<tag>/)
N=$1; TAG=$2
mkdir -p gpurun_out/$TAG
if [ "$N" = "1" ]; then TR="python"; else TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 297$N"; fi
nvidia-smi -L | wc -l > gpurun_out/$TAG/gpus_visible.txt
if [ "$N" != "1" ]; then
  $TR tools/par_check.py 8 6 4 2 > gpurun_out/$TAG/par_check_p2_n$N.txt 2>&1; tail -1 gpurun_out/$TAG/par_check_p2_n$N.txt
  $TR tools/par_check.py 6 6 4 3 > gpurun_out/$TAG/par_check_p3_n$N.txt 2>&1; tail -1 gpurun_out/$TAG/par_check_p3_n$N.txt
fi
$TR bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/$TAG/bench_a_n$N.json 2> gpurun_out/$TAG/bench_a_n$N.err
SB_BENCH_CONFIG=b $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/$TAG/bench_b_weak_n$N.json 2> gpurun_out/$TAG/bench_b_weak_n$N.err
SB_BENCH_CONFIG=b SB_BENCH_SCALING=strong $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/$TAG/bench_b_strong_n$N.json 2> gpurun_out/$TAG/bench_b_strong_n$N.err
python - <<PY
import json
for f in ['bench_a_n$N','bench_b_weak_n$N','bench_b_strong_n$N']:
    try:
        d=json.loads(open('gpurun_out/$TAG/'+f+'.json').read().strip().splitlines()[-1])
        print(f, 'GDOF/s', round(d['value'],3), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],3), 'jvp', round(d['roofline']['frac'],3), 'eval', round(d['roofline_evaluate']['frac'],3), 'parity', d.get('parity_vs_global',{}).get('ok', d.get('parity_vs_global',{}).get('skipped','')[:20] if d.get('parity_vs_global') else None))
    except Exception as e:
        print(f, 'ERR', e); print(open('gpurun_out/$TAG/'+f+'.err').read()[-1500:])
PY
