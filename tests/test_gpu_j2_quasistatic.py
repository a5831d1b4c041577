"""GPU: BASELINE.json configs[2] as ONE run -- SolidMechanics J2 plasticity with QuadratureData state, H1 p=2 hex, Newton
steps with assemble(grad) every iteration and the state committed after convergence -- against the same run done by the
reference's own kernels, material and mfem (tests/golden/newton_j2_*.npz, oracle/make_golden_krylov.py:
make_newton_j2_case, modelled on src/serac/physics/tests/solid.cpp:28-113: the beam loaded into the plastic range and
unloaded in 10 quasi-static steps).  Each step: warmStartDisplacement (solid_mechanics.hpp:1727-1771), Newton on the
residual with the essential rows zeroed (:1120-1140), then residual with updateQdata(true) (:1278-1288)."""
import numpy as np
import pytest

from harness import Product, load_golden, rel_l2

pytestmark = pytest.mark.gpu


def test_j2_quasistatic_load_unload_matches_reference_stack():
    import torch
    from serac_b200 import capi
    cuda = torch.device("cuda:0")
    g = load_golden("newton_j2_solid_j2smallstrain_linear_hex_p2_beam")
    prob = dict(g)
    prob["qdata0"] = g["qdata_initial"]
    P = Product(prob, cuda)
    assert P.integral.table.state_bytes == 80  # J2SmallStrain::State: plastic strain (9) + accumulated plastic strain
    n, ess, tip_z = P.G.LSize(), g["ess"], g["tip_z"]
    t64 = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=cuda)
    S, a = torch.zeros_like(P.S), torch.zeros(n, dtype=torch.float64, device=cuda)
    u = torch.zeros(n, dtype=torch.float64, device=cuda)
    cg = capi.CGSolver(n, rel_tol=float(g["rel_tol"]), max_iter=int(g["max_iter"]))
    num_steps = int(g["num_steps"])
    newton_its, eqps_max = [], []
    for step in range(1, num_steps + 1):
        t = step / num_steps
        # --- warm start: the new essential values through one linear solve with the current Jacobian ---
        prescribed = np.zeros(n)
        prescribed[tip_z] = t * (t - 1.0)
        dub = torch.zeros(n, dtype=torch.float64, device=cuda)
        ess_t = torch.as_tensor(ess.astype(np.int64), device=cuda)
        dub[ess_t] = t64(prescribed[ess]) - u[ess_t]
        r0 = P.F(t, [S, u, a], wrt=1)
        pat, vals = P.F.assemble(1)
        Ae = capi.eliminate_rows_cols(pat, vals, ess)
        rhs = -r0
        capi.eliminate_bc(pat, vals, Ae, ess, dub, rhs)
        capi.copy_subvector(rhs, dub, ess)
        dinv = capi.jacobi_setup(capi.csr_diagonal(pat, vals))
        dus, info = cg.Mult(pat, rhs, x=dub.clone(), values=vals, dinv=dinv, iterative_mode=True)
        assert info["converged"], (step, info)
        u = u + dus
        # --- Newton, every Jacobian assembled and eliminated ---
        norms = []
        for _ in range(int(g["newton_max"])):
            r = P.F(t, [S, u, a], wrt=1)
            capi.set_subvector(r, ess, 0.0)
            norms.append(float(torch.linalg.norm(r)))
            if norms[-1] < max(float(g["newton_rel"]) * norms[0], float(g["newton_abs"])):
                break
            pat, vals = P.F.assemble(1)
            capi.eliminate_rows_cols(pat, vals, ess, want_Ae=False)
            dinv = capi.jacobi_setup(capi.csr_diagonal(pat, vals))
            du, info = cg.Mult(pat, r, values=vals, dinv=dinv)
            assert info["converged"], (step, info)
            u = u - du
        assert norms[-1] < max(float(g["newton_rel"]) * norms[0], float(g["newton_abs"])), (step, norms)
        newton_its.append(len(norms) - 1)
        # --- equilibrium found: commit the quadrature-point state (updateQdata(true)); the residual is the reactions ---
        reactions = P.F(t, [S, u, a], update_state=True)
        torch.cuda.synchronize()
        assert rel_l2(u.cpu().numpy(), g["u_steps"][step - 1]) < 1e-8, step
        eqps_max.append(float(P.qdata.view(-1, 10)[:, 9].max()))
    assert newton_its == [int(i) for i in g["newton_iterations"]], (newton_its, g["newton_iterations"])
    assert sum(newton_its) >= 8 and max(eqps_max) > 1e-3  # the run really went plastic and assembled repeatedly
    assert np.allclose(eqps_max, g["eqps_max"], rtol=1e-7, atol=1e-12)
    assert rel_l2(P.qdata.cpu().numpy(), g["qdata_final"]) < 1e-8
    assert rel_l2(reactions.cpu().numpy(), g["reactions_final"]) < 1e-7
    # unloaded to zero tip displacement, yet not back to the undeformed beam: plastic strain stays (solid.cpp:108-112)
    assert float(torch.linalg.norm(u)) > 1e-2 and abs(float(u[int(tip_z[0])])) < 1e-14
