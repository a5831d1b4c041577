"""Times the Newton step's linear algebra (include/serac_b200_solvers.h) on config A (BASELINE.json configs[1]): the
eliminated NeoHookean Jacobian of the 32 768-hex beam (840 k rows, ~1.6e8 nonzeros), against the kernels' rooflines
(DESIGN.md section 4b): SpMV 12 B per nonzero + 16 B per row, CG update 64 B per unknown.  Prints one JSON object.

    python tools/solver_timing.py [--json out.json] [--iters 50]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
from serac_b200 import capi

dev = torch.device("cuda:0")
P = bench.build_problem(dev)
F, G = P["F"], P["G"]
t64 = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
U, A, S, dU = (t64(P["host"][k]) for k in ("U", "A", "S", "dU"))
n = G.LSize()
iters = int(sys.argv[sys.argv.index("--iters") + 1]) if "--iters" in sys.argv else 50
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs") if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else None


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


F(0.0, [S, U, A], wrt=1)
pat, vals = F.assemble(1)
nnz = int(pat.nnz)
# essential dofs: every dof of the nodes in the first 1/64 of the numbering (a clamped end, roughly)
ess = np.arange(0, 3 * (n // 3 // 64), dtype=np.int32)
elim_ms = timed(lambda: capi.eliminate_rows_cols(pat, vals, ess, want_Ae=False), reps=3, warm=1)
diag = capi.csr_diagonal(pat, vals)
dinv = capi.jacobi_setup(diag)
y = torch.empty(n, dtype=torch.float64, device=dev)
spmv_ms = timed(lambda: capi.csr_spmv(pat, vals, dU, out=y))
spmv_bytes = 12.0 * nnz + 16.0 * n
op = capi.GradientOperator(F, 1, ess)
mf_ms = timed(lambda: op.Mult(dU, out=y))
# matrix-free and assembled operators agree
y_csr = capi.csr_spmv(pat, vals, dU).clone()
y_mf = op.Mult(dU).clone()
agree = float(torch.linalg.norm(y_csr - y_mf) / torch.linalg.norm(y_csr))

b = capi.csr_spmv(pat, vals, dU).clone()  # a right-hand side with a known solution


def solve(solver, operator, **kw):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    capi.lib().sb_launch_count_reset()
    e0.record()
    x, info = solver.Mult(operator, b, dinv=dinv, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return dict(info, ms=ms, ms_per_iteration=ms / max(info["final_iter"], 1), launches=int(capi.lib().sb_launch_count()),
                error_vs_known_solution=float(torch.linalg.norm(x - dU) / torch.linalg.norm(dU)))


out = dict(config="config A: 32768 hexes, H1 p=2, NeoHookean, eliminated Jacobian", rows=int(n), nnz=nnz,
           essential_dofs=int(len(ess)), hbm_peak_gbs=peak, eliminate_rows_cols_ms=elim_ms,
           spmv_ms=spmv_ms, spmv_gbs=spmv_bytes / (spmv_ms * 1e-3) / 1e9,
           spmv_frac_of_peak=(spmv_bytes / (spmv_ms * 1e-3) / 1e9 / peak) if peak else None,
           matrix_free_operator_ms=mf_ms, matrix_free_vs_assembled_rel=agree)
for check_every in (1, 8):
    cg = capi.CGSolver(n, rel_tol=1e-8, max_iter=iters, check_every=check_every)
    out[f"cg_csr_check_every_{check_every}"] = solve(cg, pat, values=vals)
cg = capi.CGSolver(n, rel_tol=1e-8, max_iter=iters, check_every=8)
out["cg_matrix_free"] = solve(cg, op)
gm = capi.GMRESSolver(n, kdim=50, rel_tol=1e-8, max_iter=iters)
out["gmres50_csr"] = solve(gm, pat, values=vals)
print(json.dumps(out, indent=1))
if "--json" in sys.argv:
    json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
