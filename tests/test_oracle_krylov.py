"""CPU: the numpy restatement of the Newton-step linear algebra (oracle/krylov_oracle.py) against the reference's own
mfem (tests/golden/krylov_*.npz, written by oracle/make_golden_krylov.py; live against oracle/_ref when it is built)."""
import numpy as np
import pytest

from harness import krylov_golden_names, load_golden, rel_l2
from oracle import krylov_oracle as ko


@pytest.fixture(scope="module", params=krylov_golden_names())
def case(request):
    g = load_golden(request.param)
    src = load_golden(g["source"])
    g["row_ptr"], g["col_ind"], g["values"] = src["row_ptr"].astype(np.int64), src["col_ind"].astype(np.int64), src["values"]
    return g


def test_fixtures_exist():
    assert len(krylov_golden_names()) >= 4


def test_elimination_is_exact(case):
    """pure selection / subtraction of 0 or 1: bit-exact"""
    A, Ae = ko.eliminate_rows_cols(case["row_ptr"], case["col_ind"], case["values"], case["ess"])
    assert np.array_equal(A, case["values_elim"])
    assert np.array_equal(Ae, case["values_e"])
    assert np.array_equal(A + Ae, case["values"])  # A_before = A_after + Ae entry by entry


def test_eliminate_bc_and_spmv(case):
    B = ko.eliminate_bc(case["row_ptr"], case["col_ind"], case["values_elim"], case["values_e"], case["ess"], case["X"],
                        case["B"])
    assert rel_l2(B, case["B_bc"]) < 1e-14
    assert np.array_equal(B[case["ess"]], case["X"][case["ess"]])  # diagonal is exactly 1 there
    assert rel_l2(ko.spmv(case["row_ptr"], case["col_ind"], case["values_elim"], case["V"]), case["AV"]) < 1e-14


def _solve(case, **kw):
    rp, ci, A = case["row_ptr"], case["col_ind"], case["values_elim"]
    dinv = ko.jacobi_dinv(ko.diagonal(rp, ci, A)) if case["jacobi"] else None
    return ko.cg(lambda v: ko.spmv(rp, ci, A, v), case["B_bc"], dinv=dinv, rel_tol=case["rel_tol"],
                 abs_tol=case["abs_tol"], **kw)


def test_cg_matches_mfem(case):
    x, info = _solve(case, max_iter=case["max_iter"])
    assert info["converged"] and case["cg_converged"]
    assert abs(info["final_iter"] - case["cg_final_iter"]) <= 1  # summation order may move the last step
    assert abs(info["initial_norm"] / case["cg_initial_norm"] - 1) < 1e-12
    assert rel_l2(x, case["x"]) < 1e-7
    # the stopping rule itself: (B r, r) <= rel_tol^2 (B r0, r0)
    assert info["final_norm"] <= case["rel_tol"] * info["initial_norm"] * (1 + 1e-12)


def test_cg_max_iter_exit(case):
    """solvers.cpp:842-845: not converged, final_iter = max_iter, x = the 5th iterate"""
    x, info = _solve(case, max_iter=5)
    assert not info["converged"] and info["final_iter"] == 5 == case["cg5_final_iter"] and not case["cg5_converged"]
    assert rel_l2(x, case["x_maxiter5"]) < 1e-12
    assert abs(info["final_norm"] / case["cg5_final_norm"] - 1) < 1e-10


def test_cg_iterative_mode(case):
    x, info = _solve(case, max_iter=case["max_iter"], x0=case["X"])
    assert info["converged"] and abs(info["final_iter"] - case["cgg_final_iter"]) <= 1
    assert rel_l2(x, case["x_guess"]) < 1e-7


def test_live_reference_agrees_with_fixture(case):
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    A, Ae = ref.eliminate_rows_cols(case["row_ptr"], case["col_ind"], case["values"], case["ess"])
    assert np.array_equal(A, case["values_elim"]) and np.array_equal(Ae, case["values_e"])
    x, info = ref.cg(case["row_ptr"], case["col_ind"], A, case["B_bc"], jacobi=bool(case["jacobi"]),
                     rel_tol=case["rel_tol"], max_iter=case["max_iter"])
    assert info["final_iter"] == case["cg_final_iter"] and np.array_equal(x, case["x"])


def _gmres(case, **kw):
    rp, ci, A = case["row_ptr"], case["col_ind"], case["values_elim"]
    dinv = ko.jacobi_dinv(ko.diagonal(rp, ci, A)) if case["jacobi"] else None
    return ko.gmres(lambda v: ko.spmv(rp, ci, A, v), case["B_bc"], dinv=dinv, rel_tol=case["rel_tol"],
                    abs_tol=case["abs_tol"], **kw)


@pytest.mark.parametrize("kdim", [20, 50])
def test_gmres_matches_mfem(case, kdim):
    """restarted GMRES (solvers.cpp:981-1165): restarts (kdim 20 < iterations) and mfem's default kdim 50"""
    x, info = _gmres(case, kdim=kdim, max_iter=case["max_iter"])
    assert info["converged"] and case[f"gmres{kdim}_converged"]
    assert abs(info["final_iter"] - case[f"gmres{kdim}_final_iter"]) <= 1
    assert abs(info["initial_norm"] / case[f"gmres{kdim}_initial_norm"] - 1) < 1e-12
    assert rel_l2(x, case[f"x_gmres{kdim}"]) < 1e-7


def test_gmres_exits_like_mfem(case):
    """max_iter in the middle of the second cycle (kdim 5, max_iter 7): the true residual norm is reported, not converged
    (:1117-1142); and iterative_mode"""
    x, info = _gmres(case, kdim=5, max_iter=7)
    assert not info["converged"] and info["final_iter"] == 7 == case["gmres5m7_final_iter"]
    assert rel_l2(x, case["x_gmres5_maxiter7"]) < 1e-11
    assert abs(info["final_norm"] / case["gmres5m7_final_norm"] - 1) < 1e-10
    xg, ig = _gmres(case, kdim=20, max_iter=case["max_iter"], x0=case["X"])
    assert ig["converged"] and abs(ig["final_iter"] - case["gmres20g_final_iter"]) <= 1
    assert rel_l2(xg, case["x_gmres20_guess"]) < 1e-7


def test_warm_start_fixture_restated():
    """SolidMechanics::warmStartDisplacement (solid_mechanics.hpp:1727-1771) on the reference's Jacobian: EliminateBC,
    the essential rows, and the GMRES solve restated in numpy reproduce the fixture"""
    g = load_golden("warmstart_solid_neohookean_hex_p2_patch")
    src = load_golden(g["source"])
    rp, ci = src["row_ptr"].astype(np.int64), src["col_ind"].astype(np.int64)
    ess, n = g["ess"], len(g["u0"])
    du0 = np.zeros(n)
    du0[ess] = g["prescribed"] - g["u0"][ess]
    rhs = ko.eliminate_bc(rp, ci, g["values_elim"], g["values_e"], ess, du0, -g["r"])
    rhs[ess] = du0[ess]
    assert rel_l2(rhs, g["rhs"]) < 1e-14
    A = g["values_elim"]
    dinv = ko.jacobi_dinv(ko.diagonal(rp, ci, A))
    du, info = ko.gmres(lambda v: ko.spmv(rp, ci, A, v), rhs, dinv=dinv, kdim=50, rel_tol=g["rel_tol"],
                        max_iter=g["max_iter"], x0=du0)
    assert info["converged"] and abs(info["final_iter"] - g["gmres_iterations"]) <= 1
    assert rel_l2(du, g["du"]) < 1e-7
    assert np.array_equal(du[ess], du0[ess]) or rel_l2(du[ess], du0[ess]) < 1e-12  # the prescribed increments
    assert g["residual_after"] < 1e-3 * g["residual_before"]  # one linearized step: second-order small


def test_elimination_edge_cases_against_live_mfem(case):
    """no essential dofs (nothing changes), every dof essential (A = I, Ae = A - I, CG converges in one step): the numpy
    restatement against the reference's mfem"""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    rp, ci, V = case["row_ptr"], case["col_ind"], case["values"]
    n = len(rp) - 1
    for ess in (np.zeros(0, dtype=np.int32), np.arange(n, dtype=np.int32), np.array([n - 1], dtype=np.int32)):
        A, Ae = ko.eliminate_rows_cols(rp, ci, V, ess)
        A_ref, Ae_ref = ref.eliminate_rows_cols(rp, ci, V, ess)
        assert np.array_equal(A, A_ref) and np.array_equal(Ae, Ae_ref)
    A, _ = ko.eliminate_rows_cols(rp, ci, V, np.arange(n))
    assert np.array_equal(ko.diagonal(rp, ci, A), np.ones(n)) and np.count_nonzero(A) == n
    x, info = ko.cg(lambda v: ko.spmv(rp, ci, A, v), case["B"], rel_tol=1e-12, max_iter=10)
    x_ref, info_ref = ref.cg(rp, ci, A, case["B"], jacobi=False, rel_tol=1e-12, max_iter=10)
    assert info["converged"] and info["final_iter"] == info_ref["final_iter"] == 1 and np.array_equal(x, x_ref)
    # zero right-hand side: converged at iteration 0 with x = 0 (solvers.cpp:769-775, :1032-1039)
    for solve, solve_ref in ((ko.cg, ref.cg), (ko.gmres, ref.gmres)):
        x, info = solve(lambda v: ko.spmv(rp, ci, A, v), np.zeros(n), rel_tol=1e-12, max_iter=10)
        x_ref, info_ref = solve_ref(rp, ci, A, np.zeros(n), jacobi=False, rel_tol=1e-12, max_iter=10)
        assert info["converged"] and info_ref["converged"] and info["final_iter"] == info_ref["final_iter"] == 0
        assert not x.any() and not x_ref.any()
