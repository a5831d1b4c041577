#!/bin/bash
# One gpurun call that measures the solver layer: timings + rooflines on config A, the
# FP64 peak, an ncu launch list and one full capture of the SpMV + dot kernel.  Outputs under gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/measure_solvers.sh'
set -u
mkdir -p gpurun_out
python tools/solver_timing.py --iters 100 --json gpurun_out/solver_timing.json > gpurun_out/solver_timing.log 2>&1
nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/fp64_peak.cu -o tools/fp64_peak 2>/dev/null
tools/fp64_peak > gpurun_out/fp64_peak.jsonl 2>&1
# launch list of a short solve (per-launch times are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/solver_launches.csv \
    python tools/solver_timing.py --iters 10 > gpurun_out/solver_launches.log 2>&1
# one full capture of the dominant kernel
ncu --set full --clock-control none --import-source on -k regex:cg_spmv_dot_kernel -c 1 -o gpurun_out/solver_spmv_full \
    python tools/solver_timing.py --iters 10 > gpurun_out/solver_spmv_full.log 2>&1
ncu -i gpurun_out/solver_spmv_full.ncu-rep --page raw --csv > gpurun_out/solver_spmv_full_raw.csv 2>/dev/null
tail -5 gpurun_out/solver_timing.log
cat gpurun_out/fp64_peak.jsonl
