"""-m gpu: the Newton step's linear algebra on the device (include/serac_b200_solvers.h, serac_b200/csrc/krylov.cu) against
the reference's own mfem (tests/golden/krylov_*.npz: SparseMatrix::EliminateRowCol, EliminateBC, CGSolver +
OperatorJacobiSmoother on Jacobians the reference Functional assembled) and the numpy restatement (oracle/krylov_oracle).

Tolerances: bit-exact for the elimination (selection and subtraction of 0 / 1) and the diagonal; <= 1e-12 relative L2 for
products (summation order differs); CG iterates are compared after convergence: the solution within 1e-7 (both sides stop
at rel_tol 1e-8 ... 1e-10 of the initial preconditioned residual), the iteration count within 2, and the exit semantics
(max_iter, iterative_mode, converged flag) exactly."""
import numpy as np
import pytest

from harness import Product, krylov_golden_names, load_golden, rel_l2
from oracle import krylov_oracle as ko

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=krylov_golden_names())
def case(request, cuda):
    import torch
    from serac_b200 import capi
    g = load_golden(request.param)
    prob = load_golden(g["source"])
    P = Product(prob, cuda)
    pat = capi.CsrPattern(P.G, P.G)
    rp, ci = pat.to_host()
    assert np.array_equal(rp, prob["row_ptr"]) and np.array_equal(ci, prob["col_ind"])
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=cuda)
    return dict(g=g, prob=prob, P=P, pat=pat, t=t, capi=capi, torch=torch)


def _eliminated(c):
    vals = c["t"](c["prob"]["values"]).clone()
    Ae = c["capi"].eliminate_rows_cols(c["pat"], vals, c["g"]["ess"])
    return vals, Ae


def _dinv(c, vals):
    return c["capi"].jacobi_setup(c["capi"].csr_diagonal(c["pat"], vals)) if c["g"]["jacobi"] else None


def test_cpp_self_test_of_the_solver_abi(cuda):
    """tests/cpp/krylov_test.cu: the solver entry points driven from C++ on a chain-of-elements matrix, every result
    checked against host loops in that file (elimination bit-exact, CG vs a restated mfem CG)"""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp", "krylov_test")
    assert os.path.exists(exe), "tests/cpp/krylov_test is not built (python __graft_entry__.py)"
    for ne in ("3000", "200000"):
        run = subprocess.run([exe, ne], capture_output=True, text=True, timeout=300)
        assert run.returncode == 0 and "ALL PASSED" in run.stdout, run.stdout[-3000:] + run.stderr[-2000:]


def test_set_subvector(case):
    c = case
    v = c["t"](c["g"]["B"]).clone()
    c["capi"].set_subvector(v, c["g"]["ess"], 0.0)
    ref = c["g"]["B"].copy()
    ref[c["g"]["ess"]] = 0.0
    assert np.array_equal(v.cpu().numpy(), ref)


def test_eliminate_rows_cols_bit_exact(case):
    c = case
    vals, Ae = _eliminated(c)
    assert np.array_equal(vals.cpu().numpy(), c["g"]["values_elim"])
    assert np.array_equal(Ae.cpu().numpy(), c["g"]["values_e"])
    # without the eliminated part, and with an empty list (nothing changes)
    v2 = c["t"](c["prob"]["values"]).clone()
    assert c["capi"].eliminate_rows_cols(c["pat"], v2, c["g"]["ess"], want_Ae=False) is None
    assert np.array_equal(v2.cpu().numpy(), c["g"]["values_elim"])
    v3 = c["t"](c["prob"]["values"]).clone()
    Ae3 = c["capi"].eliminate_rows_cols(c["pat"], v3, np.zeros(0, dtype=np.int32))
    assert np.array_equal(v3.cpu().numpy(), c["prob"]["values"]) and not Ae3.any()


def test_diagonal_spmv_and_eliminate_bc(case):
    c, g = case, case["g"]
    vals, Ae = _eliminated(c)
    rp, ci = c["prob"]["row_ptr"].astype(np.int64), c["prob"]["col_ind"].astype(np.int64)
    assert np.array_equal(c["capi"].csr_diagonal(c["pat"], vals).cpu().numpy(), ko.diagonal(rp, ci, g["values_elim"]))
    assert rel_l2(c["capi"].csr_spmv(c["pat"], vals, c["t"](g["V"])).cpu().numpy(), g["AV"]) < 1e-12
    B = c["t"](g["B"]).clone()
    c["capi"].eliminate_bc(c["pat"], vals, Ae, g["ess"], c["t"](g["X"]), B)
    assert rel_l2(B.cpu().numpy(), g["B_bc"]) < 1e-12
    assert np.array_equal(B.cpu().numpy()[g["ess"]], g["X"][g["ess"]])


def test_cg_matches_mfem(case):
    c, g = case, case["g"]
    vals, _ = _eliminated(c)
    dinv = _dinv(c, vals)
    cg = c["capi"].CGSolver(c["pat"].num_rows, rel_tol=g["rel_tol"], abs_tol=g["abs_tol"], max_iter=g["max_iter"])
    x, info = cg.Mult(c["pat"], c["t"](g["B_bc"]), values=vals, dinv=dinv)
    assert info["converged"] and g["cg_converged"]
    assert abs(info["final_iter"] - g["cg_final_iter"]) <= 2
    assert abs(info["initial_norm"] / g["cg_initial_norm"] - 1) < 1e-12
    assert info["final_norm"] <= g["rel_tol"] * info["initial_norm"] * (1 + 1e-12)
    assert rel_l2(x.cpu().numpy(), g["x"]) < 1e-7
    # the answer solves the system: ||A x - b|| small against ||b||
    res = c["capi"].csr_spmv(c["pat"], vals, x).cpu().numpy() - g["B_bc"]
    assert np.linalg.norm(res) < 1e-6 * np.linalg.norm(g["B_bc"])
    # deterministic reductions: a second solve reproduces the first bit for bit
    x2, info2 = cg.Mult(c["pat"], c["t"](g["B_bc"]), values=vals, dinv=dinv)
    assert info2 == info and c["torch"].equal(x, x2)
    # the host may run ahead of the device-side verdict: same iterates, same count
    cg7 = c["capi"].CGSolver(c["pat"].num_rows, rel_tol=g["rel_tol"], abs_tol=g["abs_tol"], max_iter=g["max_iter"],
                             check_every=7)
    x7, info7 = cg7.Mult(c["pat"], c["t"](g["B_bc"]), values=vals, dinv=dinv)
    assert info7 == info and c["torch"].equal(x, x7)


def test_cg_exits_like_mfem(case):
    c, g = case, case["g"]
    vals, _ = _eliminated(c)
    dinv = _dinv(c, vals)
    n = c["pat"].num_rows
    # max_iter reached (solvers.cpp:842-845)
    cg5 = c["capi"].CGSolver(n, rel_tol=g["rel_tol"], max_iter=5)
    x5, info5 = cg5.Mult(c["pat"], c["t"](g["B_bc"]), values=vals, dinv=dinv)
    assert not info5["converged"] and info5["final_iter"] == 5 == g["cg5_final_iter"]
    assert rel_l2(x5.cpu().numpy(), g["x_maxiter5"]) < 1e-10
    assert abs(info5["final_norm"] / g["cg5_final_norm"] - 1) < 1e-8
    # iterative_mode: x holds the initial guess (:725-729)
    cg = c["capi"].CGSolver(n, rel_tol=g["rel_tol"], max_iter=g["max_iter"])
    xg, infog = cg.Mult(c["pat"], c["t"](g["B_bc"]), x=c["t"](g["X"]).clone(), values=vals, dinv=dinv,
                        iterative_mode=True)
    assert infog["converged"] and abs(infog["final_iter"] - g["cgg_final_iter"]) <= 2
    assert rel_l2(xg.cpu().numpy(), g["x_guess"]) < 1e-7
    # zero right-hand side: converged at once with x = 0 (:769-775)
    x0, info0 = cg.Mult(c["pat"], c["torch"].zeros(n, dtype=c["torch"].float64, device=vals.device), values=vals,
                        dinv=dinv)
    assert info0["converged"] and info0["final_iter"] == 0 and not x0.any()


@pytest.mark.parametrize("kdim", [20, 50])
def test_gmres_matches_mfem(case, kdim):
    """mfem::GMRESSolver with restarts (kdim 20) and with mfem's default restart length"""
    c, g = case, case["g"]
    vals, _ = _eliminated(c)
    dinv = _dinv(c, vals)
    gm = c["capi"].GMRESSolver(c["pat"].num_rows, kdim=kdim, rel_tol=g["rel_tol"], abs_tol=g["abs_tol"],
                               max_iter=g["max_iter"])
    x, info = gm.Mult(c["pat"], c["t"](g["B_bc"]), values=vals, dinv=dinv)
    assert info["converged"] and g[f"gmres{kdim}_converged"]
    assert abs(info["final_iter"] - g[f"gmres{kdim}_final_iter"]) <= 2
    assert abs(info["initial_norm"] / g[f"gmres{kdim}_initial_norm"] - 1) < 1e-12
    assert rel_l2(x.cpu().numpy(), g[f"x_gmres{kdim}"]) < 1e-7
    x2, info2 = gm.Mult(c["pat"], c["t"](g["B_bc"]), values=vals, dinv=dinv)
    assert info2 == info and c["torch"].equal(x, x2)  # deterministic


def test_gmres_exits_like_mfem(case):
    c, g = case, case["g"]
    vals, _ = _eliminated(c)
    dinv = _dinv(c, vals)
    n = c["pat"].num_rows
    gm5 = c["capi"].GMRESSolver(n, kdim=5, rel_tol=g["rel_tol"], max_iter=7)  # stops inside the second cycle
    x, info = gm5.Mult(c["pat"], c["t"](g["B_bc"]), values=vals, dinv=dinv)
    assert not info["converged"] and info["final_iter"] == 7 == g["gmres5m7_final_iter"]
    assert rel_l2(x.cpu().numpy(), g["x_gmres5_maxiter7"]) < 1e-10
    assert abs(info["final_norm"] / g["gmres5m7_final_norm"] - 1) < 1e-8
    gm = c["capi"].GMRESSolver(n, kdim=20, rel_tol=g["rel_tol"], max_iter=g["max_iter"])
    xg, ig = gm.Mult(c["pat"], c["t"](g["B_bc"]), x=c["t"](g["X"]).clone(), values=vals, dinv=dinv, iterative_mode=True)
    assert ig["converged"] and abs(ig["final_iter"] - g["gmres20g_final_iter"]) <= 2
    assert rel_l2(xg.cpu().numpy(), g["x_gmres20_guess"]) < 1e-7
    # matrix-free operator under GMRES
    P = c["P"]
    P.F(0.0, [P.S, P.U, P.A], wrt=1)
    op = c["capi"].GradientOperator(P.F, 1, g["ess"])
    xo, io = gm.Mult(op, c["t"](g["B_bc"]), dinv=dinv)
    assert io["converged"] and rel_l2(xo.cpu().numpy(), g["x_gmres20"]) < 1e-7


def test_matrix_free_operator_equals_eliminated_matrix(case):
    """[J_ii 0; 0 I] through Gradient::Mult == the eliminated assembled matrix (check_gradient.hpp:34-44 with BCs)"""
    c, g, P = case, case["g"], case["P"]
    P.F(0.0, [P.S, P.U, P.A], wrt=1)  # stores the derivative at the fixture's state
    op = c["capi"].GradientOperator(P.F, 1, g["ess"])
    assert rel_l2(op.Mult(c["t"](g["V"])).cpu().numpy(), g["AV"]) < 1e-12
    cg = c["capi"].CGSolver(op.n, rel_tol=g["rel_tol"], max_iter=g["max_iter"])
    vals, _ = _eliminated(c)
    x, info = cg.Mult(op, c["t"](g["B_bc"]), dinv=_dinv(c, vals))
    assert info["converged"] and abs(info["final_iter"] - g["cg_final_iter"]) <= 2
    assert rel_l2(x.cpu().numpy(), g["x"]) < 1e-7
    # no essential dofs: the plain Jacobian action
    op0 = c["capi"].GradientOperator(P.F, 1)
    assert rel_l2(op0.Mult(P.dU).cpu().numpy(), c["prob"]["dr"]) < 1e-12


def test_newton_iteration_stays_on_the_device(cuda):
    """One quasi-static solve of SolidMechanics (solid_mechanics.hpp:1120-1140 residual / Jacobian lambdas + Newton +
    CG/Jacobi) with every step on the GPU: residual with derivative store, essential rows zeroed, assembly, elimination,
    Jacobi, CG, update -- against the same solve done by the reference's kernels and mfem
    (tests/golden/newton_*.npz, oracle/make_golden_krylov.py:make_newton_case)."""
    import torch
    from serac_b200 import capi
    g = load_golden("newton_solid_neohookean_hex_p2_patch")
    P = Product(load_golden(g["source"]), cuda)
    n = P.G.LSize()
    ess = g["ess"]
    u = torch.zeros(n, dtype=torch.float64, device=cuda)
    a = float(g["force_scale"]) * P.A
    cg = capi.CGSolver(n, rel_tol=float(g["rel_tol"]), max_iter=int(g["max_iter"]))
    norms, cg_iters = [], []
    for _ in range(int(g["newton_max"])):
        r = P.F(0.0, [P.S, u, a], wrt=1)
        capi.set_subvector(r, ess, 0.0)
        norms.append(float(torch.linalg.norm(r)))
        if norms[-1] < float(g["newton_rel"]) * norms[0]:
            break
        pat, vals = P.F.assemble(1)
        capi.eliminate_rows_cols(pat, vals, ess, want_Ae=False)
        dinv = capi.jacobi_setup(capi.csr_diagonal(pat, vals))
        du, info = cg.Mult(pat, r, values=vals, dinv=dinv)
        assert info["converged"], info
        cg_iters.append(info["final_iter"])
        u -= du
    ref_norms = g["residual_norms"]
    assert len(norms) == len(ref_norms), (norms, ref_norms)  # the same number of Newton steps
    # residual norms agree while they are above rounding level; the last one is at rounding level on both sides
    # (each linear solve stops at rel_tol: the iterates agree to ~1e-10, so a residual of 1e-8 agrees to ~1e-3 relative)
    for mine, ref in zip(norms[:-1], ref_norms[:-1]):
        assert abs(mine / ref - 1) < 1e-3, (norms, ref_norms)
    assert norms[-1] < 1e-11 * norms[0]
    assert all(abs(int(i) - int(j)) <= 2 for i, j in zip(cg_iters, g["cg_iterations"])), (cg_iters, g["cg_iterations"])
    assert rel_l2(u.cpu().numpy(), g["u"]) < 1e-8
    assert not u[torch.as_tensor(ess.astype(np.int64), device=cuda)].any()  # essential dofs never move


def export_newton_case(d):
    """inputs of `functional_test newton`: the solid fixture + the Newton fixture's essential dofs and settings"""
    import os
    from test_gpu_cpp_functional import export
    g = load_golden("newton_solid_neohookean_hex_p2_patch")
    export(load_golden(g["source"]), d)
    g["ess"].astype(np.int32).tofile(os.path.join(d, "ess.bin"))
    with open(os.path.join(d, "meta.txt"), "a") as f:
        for k in ("force_scale", "rel_tol", "max_iter", "newton_rel", "newton_max"):
            f.write(f"{k} {float(g[k])!r}\n")
    return g


def check_newton_outputs(d, g):
    from test_gpu_cpp_functional import out
    for tag in ("", "_mf", "_gmres"):  # assembled-matrix CG, matrix-free CG over Gradient::Mult, GMRES(50)
        norms, iters, u = out(d, "residual_norms" + tag), out(d, "cg_iterations" + tag), out(d, "u" + tag)
        ref_norms = g["residual_norms"]
        assert len(norms) == len(ref_norms), (tag, norms, ref_norms)
        # each linear solve stops at rel_tol: the iterates agree to ~1e-10, so a residual of 1e-8 agrees to ~1e-3 relative
        assert all(abs(a / b - 1) < 1e-3 for a, b in zip(norms[:-1], ref_norms[:-1])), (tag, norms, ref_norms)
        assert norms[-1] < 1e-11 * norms[0]
        if tag != "_gmres":  # the fixture's linear solver is CG
            assert all(abs(int(a) - int(b)) <= 2 for a, b in zip(iters, g["cg_iterations"])), (tag, iters)
        assert rel_l2(u, g["u"]) < 1e-8, tag
        assert not u[g["ess"]].any()


def test_cpp_newton_solve_on_the_device(tmp_path, cuda):
    """serac_b200/solvers.hpp (DeviceCsrMatrix, EliminateRowsCols, JacobiSmoother, CGSolver, GradientOperator) under the
    C++ Functional: the quasi-static solve of tests/golden/newton_*.npz, against the reference stack's own solve"""
    from test_gpu_cpp_functional import run
    d = str(tmp_path)
    g = export_newton_case(d)
    run("newton", d)
    check_newton_outputs(d, g)


CPP_KRYLOV_CASES = ["krylov_solid_neohookean_hex_p2_patch", "krylov_solid_neohookean_tet_p2_cart"]


def export_krylov_case(name, d):
    """inputs of `functional_test krylov`"""
    import os
    from test_gpu_cpp_functional import export
    g = load_golden(name)
    export(load_golden(g["source"]), d)
    g["ess"].astype(np.int32).tofile(os.path.join(d, "ess.bin"))
    g["B_bc"].tofile(os.path.join(d, "B_bc.bin"))
    g["X"].tofile(os.path.join(d, "X.bin"))
    with open(os.path.join(d, "meta.txt"), "a") as f:
        for k in ("rel_tol", "max_iter", "jacobi"):
            f.write(f"{k} {float(g[k])!r}\n")
    return g


def check_krylov_outputs(d, g):
    from test_gpu_cpp_functional import out
    assert rel_l2(out(d, "values_elim"), g["values_elim"]) < 1e-12 and rel_l2(out(d, "values_e"), g["values_e"]) < 1e-12
    # the structural part of the elimination is exact
    assert np.array_equal(out(d, "values_elim") == 0.0, g["values_elim"] == 0.0)

    def check(name, x_key, prefix, exact_iters=False):
        conv, iters, n0, n1 = out(d, "info_" + name)
        assert bool(conv) == bool(g[prefix + "_converged"]), name
        if exact_iters:
            assert int(iters) == int(g[prefix + "_final_iter"]), name
        else:
            assert abs(int(iters) - int(g[prefix + "_final_iter"])) <= 2, (name, iters, g[prefix + "_final_iter"])
        assert abs(n0 / g[prefix + "_initial_norm"] - 1) < 1e-10, name
        assert rel_l2(out(d, "x_" + name), g[x_key]) < (1e-10 if exact_iters else 1e-7), name

    check("cg", "x", "cg")
    check("gmres20", "x_gmres20", "gmres20")
    check("gmres50", "x_gmres50", "gmres50")
    check("gmres20_guess", "x_gmres20_guess", "gmres20g")
    check("gmres5_maxiter7", "x_gmres5_maxiter7", "gmres5m7", exact_iters=True)
    conv, iters, _, _ = out(d, "info_gmres20_mf")  # matrix-free: same system up to summation order
    assert conv and abs(int(iters) - int(g["gmres20_final_iter"])) <= 2
    assert rel_l2(out(d, "x_gmres20_mf"), g["x_gmres20"]) < 1e-7


@pytest.mark.parametrize("name", CPP_KRYLOV_CASES)
def test_cpp_linear_solvers_match_mfem(name, tmp_path, cuda):
    """serac::CGSolver / serac::GMRESSolver (serac_b200/solvers.hpp) under the C++ Functional against mfem's solvers"""
    from test_gpu_cpp_functional import run
    d = str(tmp_path)
    g = export_krylov_case(name, d)
    run("krylov", d)
    check_krylov_outputs(d, g)
