"""Host-side timeline of the e2e step (async host buffers): how long each C-ABI call blocks the host."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
dev = torch.device("cuda:0")
P = bench.build_problem(dev)
F = P["F"]
pin = lambda a: torch.as_tensor(a, dtype=torch.float64).pin_memory()
hU, hA, hS, hdU = (pin(P["host"][k]) for k in ("U", "A", "S", "dU"))
hr, hdr = (torch.empty(P["tdofs"], dtype=torch.float64).pin_memory() for _ in range(2))
nU, nA, nS, ndU, nr, ndr = (x.numpy() for x in (hU, hA, hS, hdU, hr, hdr))
for mode in (True, False):
    for _ in range(5):
        F(0.0, [nS, nU, nA], wrt=1, out=nr, async_host=mode); F.GradientMult(1, ndU, out=ndr, async_host=mode); F.wait()
    torch.cuda.synchronize()
    t = np.zeros(4)
    n = 30
    for _ in range(n):
        t0 = time.perf_counter(); F(0.0, [nS, nU, nA], wrt=1, out=nr, async_host=mode)
        t1 = time.perf_counter(); F.GradientMult(1, ndU, out=ndr, async_host=mode)
        t2 = time.perf_counter(); F.wait()
        t3 = time.perf_counter(); torch.cuda.synchronize(); t4 = time.perf_counter()
        t += [t1 - t0, t2 - t1, t3 - t2, t4 - t0]
    print("async" if mode else "sync ", "evaluate call %.0f us, jvp call %.0f us, wait %.0f us, total %.0f us" % tuple(t / n * 1e6))
