"""Determinism stress of the bench step on ONE GPU: evaluate (with derivative store) + Jacobian action, repeated, every
result compared bit for bit with the first.  Any difference is a race in a kernel (nothing here is multi-GPU).

    SB_BENCH_CONFIG=b SB_BENCH_B_N=68 python tools/repeat_check.py [repetitions]

Written for the intermittent 128-entry difference in the Jacobian action seen twice in the 8-GPU strong config-B series
(DESIGN.md section 5): partitions of 67 / 68 hexes per direction, p = 3."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench

dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
box = os.environ.get("SB_REPEAT_BOX")  # e.g. "68,68,67"
if box and bench.BENCH_CONFIG == "b":
    bench.PER_GPU_BOX = tuple(int(x) for x in box.split(","))
P = bench.build_structured_problem(dev) if bench.BENCH_CONFIG == "b" else bench.build_problem(dev)
F = P["F"]
t64 = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
U, A, S, dU = (t64(P["host"][k]) for k in ("U", "A", "S", "dU"))
graph_mode = os.environ.get("SB_REPEAT_GRAPH", "1") != "0"
r = torch.empty(P["tdofs"], dtype=torch.float64, device=dev)
dr = torch.empty_like(r)


def step():
    F(0.0, [S, U, A], wrt=1, out=r)
    F.GradientMult(1, dU, out=dr)


for _ in range(3):
    step()
torch.cuda.synchronize()
r0, dr0 = r.clone(), dr.clone()
D0 = P["integral"].export_derivatives(1).clone() if P["ne"] < 400000 else None
if graph_mode:
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        step()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=side):
            step()
    torch.cuda.synchronize()
bad = []
for i in range(reps):
    if graph_mode:
        g.replay()
    else:
        step()
    torch.cuda.synchronize()
    nr, ndr = int((r != r0).sum().item()), int((dr != dr0).sum().item())
    if nr or ndr:
        where = torch.nonzero(dr != dr0).flatten()[:8].tolist()
        bad.append((i, nr, ndr, float((dr - dr0).abs().max().item()), where))
        if D0 is not None:
            D = P["integral"].export_derivatives(1)
            bad[-1] += (int((D != D0).sum().item()),)
print(f"{P['mesh']}: {reps} repetitions ({'graph replay' if graph_mode else 'kernel by kernel'}), "
      f"{len(bad)} differed from the first")
for b in bad[:20]:
    print("  rep, r entries, dr entries, max |diff| in dr, first indices[, derivative entries]:", b)
