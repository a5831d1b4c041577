"""Times assemble(Gradient&) on config A (BASELINE.json configs[1]/[2]): sparsity pattern creation (once), element
stiffness kernel, CSR assembly -- the FP64-bound phase of SURVEY.md §8d.  Prints one JSON object."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from serac_b200 import capi

dev = torch.device("cuda:0")
P = bench.build_problem(dev)
F, G, I = P["F"], P["G"], P["integral"]
t64 = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
U, A, S, dU = (t64(P["host"][k]) for k in ("U", "A", "S", "dU"))
F(0.0, [S, U, A], wrt=1)


def timed(fn, reps=5, warm=2):
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


import time
t0 = time.perf_counter()
pat, vals = F.assemble(1)  # first call builds the pattern (GradientAssemblyLookupTables::init) and the assembly plan
torch.cuda.synchronize()
first_ms = (time.perf_counter() - t0) * 1e3
assemble_ms = timed(lambda: F.assemble(1), reps=10)
if os.environ.get("SB_ASM_QUICK"):
    print(json.dumps(dict(warps=os.environ.get("SB_ASM_WARPS", "default"), assemble_ms=assemble_ms, first_ms=first_ms)))
    sys.exit(0)
npe = 27 * 3
K_e = torch.zeros((P["ne"], npe, npe), dtype=torch.float64, device=dev)


def elem_grad():
    K_e.zero_()
    I.ComputeElementGradients(1, K_e)


# the generic route, for comparison: element matrices in HBM (the reference's scheme), then an ordered gather per nonzero
eg_ms = timed(elem_grad)
zero_ms = timed(lambda: K_e.zero_())
elem_grad()
vals_ref = torch.empty(pat.nnz, dtype=torch.float64, device=dev)
csr_ref_ms = timed(lambda: pat.assemble(K_e, vals_ref))
rel = float((vals_ref - vals).norm() / vals_ref.norm())
assert rel < 1e-13, rel
import numpy as np
import scipy.sparse as sp
rp, ci = pat.to_host()
Acsr = sp.csr_matrix((vals.cpu().numpy(), ci, rp), shape=(G.LSize(), G.LSize()))
jvp = F.GradientMult(1, dU).cpu().numpy()
err = float(np.linalg.norm(Acsr @ dU.cpu().numpy() - jvp) / np.linalg.norm(jvp))
# FP64 roofline of the assembly.  Per (element, test node) unit and component pair: 27 points x
# (9 FMA contraction with the test gradient + 9 FMA stage z), then 9 threads x 27 FMA (stage y) and 9 x 18 (stage x)
fma_per_unit = 9 * (27 * 18 + 9 * 27 + 9 * 18)
flops = 2.0 * P["ne"] * 27 * fma_per_unit
fp64_peak_tflops = 37.1  # profiles/r2a_fp64_peak_b200.json (tools/fp64_peak.cu on this pool's B200)
hbm_bytes = P["ne"] * 27 * 8.0 * I.derivative_stored_entries(1) + pat.nnz * 8.0 + P["ne"] * 729 * 2.0
hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
out = dict(config="config A: 32768 hexes, H1 p=2, NeoHookean", nnz=int(pat.nnz), rows=int(pat.num_rows),
           first_assemble_with_pattern_and_plan_ms=first_ms, assemble_ms=assemble_ms,
           generic_route=dict(element_gradient_ms=eg_ms - zero_ms, K_e_GB=K_e.numel() * 8 / 1e9, csr_gather_ms=csr_ref_ms),
           fused_vs_generic_rel=rel, matrix_vs_action_rel=err,
           roofline_assemble=dict(kernel=f"assemble_rows_kernel<3,3,3,3,3,{I.derivative_mask(1)},*>", bound="fp64", flops=flops,
                                  achieved_tflops=flops / (assemble_ms * 1e-3) / 1e12, peak_tflops=fp64_peak_tflops,
                                  frac=flops / (assemble_ms * 1e-3) / 1e12 / fp64_peak_tflops,
                                  hbm_algorithmic_bytes=hbm_bytes, hbm_achieved_gbs=hbm_bytes / (assemble_ms * 1e-3) / 1e9,
                                  hbm_frac=hbm_bytes / (assemble_ms * 1e-3) / 1e9 / hbm_peak),
           assembled_rows_per_s=pat.num_rows / (assemble_ms * 1e-3))
print(json.dumps(out, indent=1))
if "--json" in sys.argv:
    json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
