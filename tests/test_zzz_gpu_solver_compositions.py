"""-m gpu: compositions of the solver layer written after the round's GPU minutes were spent (DESIGN.md section 7,
profiles/r1f_summary.md): every entry point they call was exercised on the B200 by tests/test_zz_gpu_krylov.py and the
C++ drivers, and their expectations were checked on the CPU against the reference's kernels + mfem or the numpy oracle, but
these tests themselves run on hardware for the first time at round end -- hence a file of their own, after the verified
ones.  Tolerances as in test_zz_gpu_krylov.py."""
import numpy as np
import pytest

from harness import Product, load_golden, rel_l2
from test_zz_gpu_krylov import _dinv, _eliminated, case  # noqa: F401  (`case` is the module-scoped fixture)

pytestmark = pytest.mark.gpu


def test_l1_jacobi_scaling(case):
    """HypreSmoother::l1Jacobi (Preconditioner::HypreL1Jacobi): the row l1 norms hypre_ParCSRComputeL1Norms(A, 1) defines
    (hypre is not in the reference tree: restated, parity unpinned), and GMRES preconditioned with them reaches the
    solution mfem's Jacobi-preconditioned solve reached"""
    c, g = case, case["g"]
    vals, _ = _eliminated(c)
    l1 = c["capi"].csr_row_l1_norms(c["pat"], vals)
    rp = c["prob"]["row_ptr"].astype(np.int64)
    want = np.add.reduceat(np.abs(g["values_elim"]), rp[:-1])
    assert rel_l2(l1.cpu().numpy(), want) < 1e-14
    dinv = c["capi"].jacobi_setup(l1)
    gm = c["capi"].GMRESSolver(c["pat"].num_rows, kdim=50, rel_tol=g["rel_tol"], max_iter=g["max_iter"])
    x, info = gm.Mult(c["pat"], c["t"](g["B_bc"]), values=vals, dinv=dinv)
    assert info["converged"] and rel_l2(x.cpu().numpy(), g["x_gmres50"]) < 1e-5


@pytest.mark.parametrize("max_iter", [0, 1, 3])
def test_iteration_limits_against_live_mfem(case, max_iter):
    """the exits at tiny iteration limits (solvers.cpp:798, :842-845, :1053, :1140-1142) against the reference's mfem run
    on the box (oracle/_ref travels with the repository): same flags, same count, same iterate"""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    c, g = case, case["g"]
    vals, _ = _eliminated(c)
    dinv = _dinv(c, vals)
    n = c["pat"].num_rows
    rp, ci = c["prob"]["row_ptr"], c["prob"]["col_ind"]
    b = c["t"](g["B_bc"])
    x, info = c["capi"].CGSolver(n, rel_tol=g["rel_tol"], max_iter=max_iter).Mult(c["pat"], b, values=vals, dinv=dinv)
    x_ref, info_ref = ref.cg(rp, ci, g["values_elim"], g["B_bc"], jacobi=bool(g["jacobi"]), rel_tol=g["rel_tol"],
                             max_iter=max_iter)
    assert (info["converged"], info["final_iter"]) == (info_ref["converged"], info_ref["final_iter"])
    assert rel_l2(x.cpu().numpy(), x_ref) < 1e-12 and abs(info["final_norm"] / info_ref["final_norm"] - 1) < 1e-10
    gm = c["capi"].GMRESSolver(n, kdim=2, rel_tol=g["rel_tol"], max_iter=max_iter)
    x, info = gm.Mult(c["pat"], b, values=vals, dinv=dinv)
    x_ref, info_ref = ref.gmres(rp, ci, g["values_elim"], g["B_bc"], jacobi=bool(g["jacobi"]), kdim=2,
                                rel_tol=g["rel_tol"], max_iter=max_iter)
    assert (info["converged"], info["final_iter"]) == (info_ref["converged"], info_ref["final_iter"])
    assert abs(info["final_norm"] / info_ref["final_norm"] - 1) < 1e-10
    if max_iter == 0:
        assert not x.any() and not x_ref.any()
    else:
        assert rel_l2(x.cpu().numpy(), x_ref) < 1e-12


def test_warm_start_displacement(cuda):
    """SolidMechanics::warmStartDisplacement (solid_mechanics.hpp:1727-1771) on the device: new essential values at an
    equilibrium state -> du on the essential dofs, r = -residual, assemble + eliminate, mfem::EliminateBC, essential rows
    = the increments, GMRES(50) + Jacobi from du as the initial guess -- against the same steps done by the reference's
    kernels and mfem (tests/golden/warmstart_*.npz)."""
    import torch
    from serac_b200 import capi
    g = load_golden("warmstart_solid_neohookean_hex_p2_patch")
    P = Product(load_golden(g["source"]), cuda)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=cuda)
    n, ess = P.G.LSize(), g["ess"]
    u0, a = t(g["u0"]), float(g["force_scale"]) * P.A
    du0 = np.zeros(n)
    du0[ess] = g["prescribed"] - g["u0"][ess]
    du = t(du0)
    r = P.F(0.0, [P.S, u0, a], wrt=1)
    assert rel_l2(r.cpu().numpy(), g["r"]) < 1e-10  # the residual at equilibrium is reactions only
    pat, vals = P.F.assemble(1)
    Ae = capi.eliminate_rows_cols(pat, vals, ess)
    assert rel_l2(vals.cpu().numpy(), g["values_elim"]) < 1e-12 and rel_l2(Ae.cpu().numpy(), g["values_e"]) < 1e-12
    rhs = -r
    capi.eliminate_bc(pat, vals, Ae, ess, du, rhs)
    capi.copy_subvector(rhs, du, ess)
    assert rel_l2(rhs.cpu().numpy(), g["rhs"]) < 1e-12
    dinv = capi.jacobi_setup(capi.csr_diagonal(pat, vals))
    gm = capi.GMRESSolver(n, kdim=50, rel_tol=float(g["rel_tol"]), max_iter=int(g["max_iter"]))
    x, info = gm.Mult(pat, rhs, x=du.clone(), values=vals, dinv=dinv, iterative_mode=True)
    assert info["converged"] and abs(info["final_iter"] - int(g["gmres_iterations"])) <= 2
    assert rel_l2(x.cpu().numpy(), g["du"]) < 1e-7
    # one linearized step from equilibrium leaves a second-order residual
    r1 = P.F(0.0, [P.S, u0 + x, a])
    capi.set_subvector(r1, ess, 0.0)
    assert abs(float(torch.linalg.norm(r1)) / float(g["residual_after"]) - 1) < 1e-4


@pytest.mark.parametrize("name", ["solid_neohookean_hex_p2_patch", "solid_linear_isotropic_hex_p2_cart",
                                  "solid_neohookean_tet_p2_cart"])
def test_implicit_dynamics_jacobian(name, cuda):
    """The second-order Jacobian of SolidMechanics (solid_mechanics.hpp:1187-1207): K = dR/du and M = dR/da assembled
    into the same pattern, J = M + c0 K (mfem::Add), essential dofs eliminated -- checked through its action against the
    reference's Jacobian actions (dr_da + c0 dr of the fixture) and solved with CG (M + c0 K is symmetric positive
    definite)."""
    import torch
    from serac_b200 import capi
    prob = load_golden(name)
    P = Product(prob, cuda)
    c0 = 0.25 * 0.01 ** 2  # Newmark beta dt^2
    P.F(0.0, [P.S, P.U, P.A], wrt=1)
    pat_k, K = P.F.assemble(1)
    P.F(0.0, [P.S, P.U, P.A], wrt=2)
    pat_m, M = P.F.assemble(2)
    rp_k, ci_k = pat_k.to_host()
    rp_m, ci_m = pat_m.to_host()
    assert np.array_equal(rp_k, rp_m) and np.array_equal(ci_k, ci_m)  # same spaces, same graph
    J = M.clone()
    capi.vector_axpy(J, c0, K)  # J = M + c0 K
    want = prob["dr_da"] + c0 * prob["dr"]
    assert rel_l2(capi.csr_spmv(pat_k, J, P.dU).cpu().numpy(), want) < 1e-12
    n = P.G.LSize()
    ess = np.arange(0, n, 7, dtype=np.int32)
    capi.eliminate_rows_cols(pat_k, J, ess, want_Ae=False)
    dinv = capi.jacobi_setup(capi.csr_diagonal(pat_k, J))
    b = capi.csr_spmv(pat_k, J, P.dU)  # known solution
    x, info = capi.CGSolver(n, rel_tol=1e-12, max_iter=2000).Mult(pat_k, b, values=J, dinv=dinv)
    assert info["converged"] and rel_l2(x.cpu().numpy(), P.dU.cpu().numpy()) < 1e-8


def test_solver_layer_at_config_a_size(cuda):
    """BASELINE.json configs[1] at full size (839 619 rows, ~1.6e8 nonzeros): the solver layer through size-independent
    properties -- the eliminated assembled matrix and the matrix-free operator agree; the host running ahead of the CG
    verdict changes nothing; CG's recurrence residual is the true residual; GMRES's preconditioned residual norm never
    grows and the norm it reports at the max_iter exit is the true one."""
    import torch
    from serac_b200 import capi
    P = Product(load_golden("config_a_beam_hex_r4_p2"), cuda)
    n = P.G.LSize()
    ess = np.arange(0, 3 * (n // 3 // 64), dtype=np.int32)
    P.F(0.0, [P.S, P.U, P.A], wrt=1)
    pat, vals = P.F.assemble(1)
    capi.eliminate_rows_cols(pat, vals, ess, want_Ae=False)
    dinv = capi.jacobi_setup(capi.csr_diagonal(pat, vals))
    assert bool(torch.isfinite(dinv).all()) and bool((dinv > 0).all())  # positive diagonal
    op = capi.GradientOperator(P.F, 1, ess)
    y_csr, y_mf = capi.csr_spmv(pat, vals, P.dU), op.Mult(P.dU)
    assert rel_l2(y_mf.cpu().numpy(), y_csr.cpu().numpy()) < 1e-11
    assert torch.equal(y_csr[:len(ess)], P.dU[:len(ess)])  # identity on the essential rows
    b = y_csr.clone()  # a system with a known solution

    def b_norm(r):  # sqrt((B r, r))
        return float(np.sqrt(capi.vector_dot(r, dinv * r)))

    x1, i1 = capi.CGSolver(n, rel_tol=1e-12, max_iter=40).Mult(pat, b, values=vals, dinv=dinv)
    x8, i8 = capi.CGSolver(n, rel_tol=1e-12, max_iter=40, check_every=8).Mult(pat, b, values=vals, dinv=dinv)
    assert i1 == i8 and torch.equal(x1, x8)
    assert i1["final_iter"] == 40 or i1["converged"]
    assert i1["final_norm"] < i1["initial_norm"]
    r_true = b - capi.csr_spmv(pat, vals, x1)
    assert abs(b_norm(r_true) / i1["final_norm"] - 1) < 1e-3
    xm, im = capi.CGSolver(n, rel_tol=1e-12, max_iter=40, check_every=8).Mult(op, b, dinv=dinv)  # matrix-free
    # same Krylov space up to the operators' summation order; rounding differences grow with the iteration count
    assert im["final_iter"] == i1["final_iter"] and rel_l2(xm.cpu().numpy(), x1.cpu().numpy()) < 1e-6
    xg, ig = capi.GMRESSolver(n, kdim=20, rel_tol=1e-12, max_iter=40).Mult(pat, b, values=vals, dinv=dinv)
    assert not ig["converged"] and ig["final_iter"] == 40 and ig["final_norm"] <= ig["initial_norm"]
    rg = dinv * (b - capi.csr_spmv(pat, vals, xg))
    assert abs(float(np.sqrt(capi.vector_dot(rg, rg))) / ig["final_norm"] - 1) < 1e-9
    # both iterations move towards the known solution
    e0 = float(torch.linalg.norm(P.dU))
    assert float(torch.linalg.norm(x1 - P.dU)) < e0 and float(torch.linalg.norm(xg - P.dU)) < e0


@pytest.mark.parametrize("name", ["solid_neohookean_hex_p2_patch", "solid_neohookean_tet_p2_cart"])
def test_cpp_equation_solver_suite(name, tmp_path, cuda):
    """The reference's own known-answer test of the solver stack (src/serac/numerics/tests/equationsolver.cpp:33-121:
    Functional H1(H1) with source 0.5 sin(u), flux grad u; solve r(x) - r(x_exact) = 0 with Newton and CG / GMRES +
    Jacobi; assert |x - x_exact| / x_exact < 1e-6 per dof), ported to the C++ host side on this fixture's mesh at p = 2,
    every step on the device -- with the assembled Jacobian (CG, GMRES), matrix-free (GMRES), and with HeatTransfer's default
    combination GMRES + l1-Jacobi."""
    from test_gpu_cpp_functional import export, out, run
    d = str(tmp_path)
    export(load_golden(name), d)
    log = run("equation_solver", d)  # non-zero exit (any FAIL) raises
    assert log.count("PASS") == 4 and "FAIL" not in log, log
    xe = out(d, "x_exact")
    for tag in ("cg", "gmres", "gmres_mf", "gmres_l1"):
        assert np.max(np.abs(out(d, "x_" + tag) - xe) / xe) < 1e-6
