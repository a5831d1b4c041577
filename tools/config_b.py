"""BASELINE.json configs[4] on ONE GPU: synthetic structured hex cube, H1 p=3, NeoHookean --
size-independent parity properties at full size plus per-phase timings against the HBM roofline.

    python tools/config_b.py [N] [--json out.json]     # N hexes per direction (default 96; 135 = ~200 M vdofs)

Properties checked (the reference at this size would take hours on the host, so parity at full size is through
properties; elementwise parity of the same kernels is pinned at small sizes by tests/test_gpu_parity.py):
  * sum of the internal-force residual over all nodes vanishes per component (partition of unity of the test basis)
  * the Jacobian action is linear, symmetric (hyperelastic tangent) and equals the central finite difference of the
    residual (check_gradient.hpp:60-90)
Algorithmic bytes per quadrature point (p=3): JVP 672, residual + derivative store ~776."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from serac_b200 import capi
from serac_b200.structured import StructuredHexMesh

NEOHOOKEAN = [1.0, 100.0, 50.0, 1.0]


def build(N, device, p=3):
    mesh = StructuredHexMesh(N, N, N)
    tab, tab1 = mesh.element_table(p), mesh.element_table(1)
    G = capi.ElementRestriction(tab, 3, mesh.num_nodes(p))
    G_shape = capi.ElementRestriction(tab1, 3, mesh.num_nodes(1))
    nodes = torch.as_tensor(mesh.vertex_coordinates(), dtype=torch.float64, device=device)
    X_q, J_q = capi.geometric_factors(capi.CUBE, 1, p + 1, G_shape, nodes)
    integral = capi.Integral(f"solid_neohookean_hex_p{p}", NEOHOOKEAN, X_q, J_q, mesh.ne)
    F = capi.Functional(G, [G_shape, G, G])
    F.AddDomainIntegral(integral, [0, 1, 2])
    return dict(mesh=mesh, G=G, G_shape=G_shape, F=F, integral=integral, keep=(nodes, X_q, J_q), ne=mesh.ne,
                nq=mesh.ne * (p + 1)**3, p=p)


def fields(P, device, h):
    n, ns = P["G"].LSize(), P["G_shape"].LSize()
    i = torch.arange(n, dtype=torch.float64, device=device)
    U = 0.02 * h * torch.sin(0.37 * i + 1.0)  # small against the element size: F = I + du/dX stays invertible
    dU = 0.5 + 0.5 * torch.sin(0.11 * i)
    dV = torch.cos(0.23 * i)
    A = torch.zeros(n, dtype=torch.float64, device=device)
    S = torch.zeros(ns, dtype=torch.float64, device=device)
    return U, A, S, dU, dV


def check_properties(P, device):
    F, h = P["F"], 1.0 / P["mesh"].n[0]
    U, A, S, dU, dV = fields(P, device, h)
    r = F(0.0, [S, U, A], wrt=1)
    out = {}
    rs = r.view(-1, 3).sum(0).abs().max().item() / r.abs().max().item()
    out["sum_of_internal_forces_rel"] = rs
    Ju = F.GradientMult(1, dU).clone()
    Jv = F.GradientMult(1, dV).clone()
    lin = F.GradientMult(1, 0.3 * dU - 1.7 * dV)
    out["linearity_rel"] = ((lin - (0.3 * Ju - 1.7 * Jv)).norm() / lin.norm()).item()
    a, b = torch.dot(dV, Ju).item(), torch.dot(dU, Jv).item()
    out["symmetry_rel"] = abs(a - b) / max(abs(a), abs(b))
    eps = 1e-6 * h
    rp = F(0.0, [S, U + eps * dU, A]).clone()
    rm = F(0.0, [S, U - eps * dU, A]).clone()
    out["finite_difference_rel"] = (((rp - rm) / (2 * eps) - Ju).norm() / Ju.norm()).item()
    del rp, rm
    F(0.0, [S, U, A], wrt=1)  # restore the stored derivatives at U
    return out


def time_phases(P, device, reps=5):
    F, I, G = P["F"], P["integral"], P["G"]
    U, A, S, dU, _ = fields(P, device, 1.0 / P["mesh"].n[0])
    r_E = torch.empty(G.ESize(), dtype=torch.float64, device=device)
    L = torch.empty(G.LSize(), dtype=torch.float64, device=device)
    fl = [P["G_shape"].field(S), G.field(U), G.field(A)]

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    return dict(residual_ms=timed(lambda: I.Mult(0.0, fl, r_E)),
                residual_with_derivative_store_ms=timed(lambda: I.Mult(0.0, fl, r_E, wrt=1)),
                jvp_ms=timed(lambda: I.GradientMult(1, G.field(dU), r_E)),
                scatter_ms=timed(lambda: G.Scatter(r_E, L)),
                scatter_node_major_ms=timed(lambda: G.Scatter(r_E, L, node_major=True)),
                step_ms=timed(lambda: (F(0.0, [S, U, A], wrt=1), F.GradientMult(1, dU))))


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    N = int(args[0]) if args else 96
    p = int(sys.argv[sys.argv.index("--p") + 1]) if "--p" in sys.argv else 3
    out_path = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
    device = torch.device("cuda:0")
    t0 = time.time()
    P = build(N, device, p)
    torch.cuda.synchronize()
    setup_s = time.time() - t0
    nq, vdofs = P["nq"], P["G"].LSize()
    props = check_properties(P, device)
    ph = time_phases(P, device)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = float(peaks.get("hbm_gbs", 6450.0))
    # V = 8 * 3 p^3 / (p+1)^3 B/qpt per L-vector pass, 4 B/qpt of indices, 648 B/qpt of derivatives
    V = 24.0 * p**3 / (p + 1)**3
    jvp_B, ev_B = 2 * V + 648 + 4, 3 * V + 0.9 + 96 + 648 + 4
    jvp_gbs = jvp_B * nq / (ph["jvp_ms"] * 1e-3) / 1e9
    ev_gbs = ev_B * nq / (ph["residual_with_derivative_store_ms"] * 1e-3) / 1e9
    res = dict(config=f"structured {N}^3 hexes, H1 p={p} vdim 3, NeoHookean, 1 GPU", hexes=P["ne"], vdofs=vdofs, qpts=nq,
               stored_derivative_GB=648.0 * nq / 1e9, setup_s=setup_s, properties=props, phases=ph,
               gdofs_per_s_step=vdofs / (ph["step_ms"] * 1e-3) / 1e9, jvp_GBps=jvp_gbs, jvp_frac_of_peak=jvp_gbs / peak,
               evaluate_GBps=ev_gbs, evaluate_frac_of_peak=ev_gbs / peak, hbm_peak_GBps=peak,
               mem_allocated_GB=torch.cuda.max_memory_allocated() / 1e9)
    print(json.dumps(res, indent=1))
    if out_path:
        json.dump(res, open(out_path, "w"), indent=1)
    ok = (props["sum_of_internal_forces_rel"] < 1e-9 and props["linearity_rel"] < 1e-12 and
          props["symmetry_rel"] < 1e-10 and props["finite_difference_rel"] < 1e-6)
    print("PASS" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
