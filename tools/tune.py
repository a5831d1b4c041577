"""Times evaluate(wrt=1) and the Jacobian action of every 'tune_*' variant (make -C serac_b200/csrc TUNE=1) on config A."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from serac_b200 import capi

dev = torch.device("cuda:0")
P = bench.build_problem(dev)
G, Gs = P["G"], P["G_shape"]
nodes, X_q, J_q = P["keep"]
t64 = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
U, A, S, dU = (t64(P["host"][k]) for k in ("U", "A", "S", "dU"))
r_E = torch.empty(G.ESize(), dtype=torch.float64, device=dev)
fields = [Gs.field(S), G.field(U), G.field(A)]
names = [n for n in capi.named_integrands() if n.startswith("tune_")] + ["solid_neohookean_hex_p2"]
pat = sys.argv[1] if len(sys.argv) > 1 else ""
ref = None
for name in names:
    if pat and pat not in name: continue
    I = capi.Integral(name, bench.NEOHOOKEAN, X_q, J_q, P["ne"])
    def timeit(fn, n=20):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    t_plain = timeit(lambda: I.Mult(0.0, fields, r_E))
    t_ad = timeit(lambda: I.Mult(0.0, fields, r_E, wrt=1))
    rr = r_E.clone()
    t_jvp = timeit(lambda: I.GradientMult(1, G.field(dU), r_E))
    if ref is None: ref = (rr.clone(), r_E.clone())
    e1 = float((rr - ref[0]).norm() / ref[0].norm()); e2 = float((r_E - ref[1]).norm() / ref[1].norm())
    print(f"{name:28s} plain {t_plain:7.1f} us   ad {t_ad:7.1f} us   jvp {t_jvp:7.1f} us   (diff vs first: {e1:.1e} {e2:.1e})", flush=True)
    del I
