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


e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
pat, vals = F.assemble(1)  # first call builds the pattern (GradientAssemblyLookupTables::init)
e1.record()
torch.cuda.synchronize()
first_ms = e0.elapsed_time(e1)
assemble_ms = timed(lambda: F.assemble(1))
npe = 27 * 3
K_e = torch.zeros((P["ne"], npe, npe), dtype=torch.float64, device=dev)


def elem_grad():
    K_e.zero_()
    I.ComputeElementGradients(1, K_e)


eg_ms = timed(elem_grad)
zero_ms = timed(lambda: K_e.zero_())
elem_grad()
ref_pat = capi.CsrPattern(G, G)  # gather lists for the reference K_e layout (what the standalone kernel wrote)
vals_ref = torch.empty(ref_pat.nnz, dtype=torch.float64, device=dev)
csr_ref_ms = timed(lambda: ref_pat.assemble(K_e, vals_ref))
assert torch.equal(vals_ref, vals), "node-major and reference-layout assembly must agree bit for bit"
# SpMV check against the matrix-free action
import numpy as np
import scipy.sparse as sp
rp, ci = pat.to_host()
Acsr = sp.csr_matrix((vals.cpu().numpy(), ci, rp), shape=(G.LSize(), G.LSize()))
jvp = F.GradientMult(1, dU).cpu().numpy()
err = float(np.linalg.norm(Acsr @ dU.cpu().numpy() - jvp) / np.linalg.norm(jvp))
flops = P["ne"] * 27 * 81 * (81 * 2 + 4 * 81 * 2)  # per element: 27 trial nodes x (81 FMA products + integrate 9 comps) -- order of magnitude
out = dict(config="config A: 32768 hexes, H1 p=2, NeoHookean", nnz=int(pat.nnz), rows=int(pat.num_rows),
           first_assemble_with_pattern_ms=first_ms, assemble_ms=assemble_ms, element_gradient_ms=eg_ms - zero_ms,
           K_e_GB=K_e.numel() * 8 / 1e9, csr_assemble_reference_layout_ms=csr_ref_ms,
           csr_assemble_node_major_ms=assemble_ms - (eg_ms - zero_ms), matrix_vs_action_rel=err,
           assembled_rows_per_s=pat.num_rows / (assemble_ms * 1e-3))
print(json.dumps(out, indent=1))
if "--json" in sys.argv:
    json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
