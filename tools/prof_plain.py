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
for name in sys.argv[1:]:
    I = capi.Integral(name, bench.NEOHOOKEAN, X_q, J_q, P["ne"])
    for _ in range(2):
        I.Mult(0.0, fields, r_E)
    for _ in range(2):
        I.Mult(0.0, fields, r_E, wrt=1)
    for _ in range(2):
        I.GradientMult(1, G.field(dU), r_E)
    torch.cuda.synchronize()
