"""-m gpu: what the adjoint / sensitivity callers of SolidMechanics need from the assembled Jacobians
(src/serac/physics/solid_mechanics.hpp:1362-1371 reverseAdjointTimestep: J^T = assemble(drdu)->Transpose(), essential
rows / columns eliminated, linear solve; :1404-1424 computeTimestep[Shape]Sensitivity: assemble(drdshape)->MultTranspose).

sb_csr_transpose / sb_csr_transpose_values against scipy (pattern bit-exact, values an exact permutation), the transposed
action, the adjoint solve against a direct solve, and the adjoint identity  lambda . (dr/dp dp) = (dr/dp^T lambda) . dp  with
the matrix-free Jacobian action on the left."""
import numpy as np
import pytest

from harness import Product, load_golden, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P(cuda):
    prob = load_golden("solid_neohookean_hex_p2_patch")
    return prob, Product(prob, cuda)


def _scipy(pat, vals, ncols):
    import scipy.sparse as sp
    rp, ci = pat.to_host()
    return sp.csr_matrix((vals.cpu().numpy(), ci, rp), shape=(len(rp) - 1, ncols))


def test_transpose_square_and_rectangular(P):
    prob, Q = P
    for wrt, ncols in ((1, len(prob["U"])), (0, len(prob["S"]))):
        Q.F(0.0, [Q.S, Q.U, Q.A], wrt=wrt)
        pat, vals = Q.F.assemble(wrt)
        A = _scipy(pat, vals, ncols)
        patT = pat.transpose()
        valsT = patT.transpose_values(vals)
        rpT, ciT = patT.to_host()
        At = A.T.tocsr()
        At.sort_indices()
        assert patT.num_rows == ncols and patT.nnz == pat.nnz
        assert np.array_equal(rpT, At.indptr) and np.array_equal(ciT, At.indices)
        assert np.array_equal(valsT.cpu().numpy(), At.data)  # a permutation of the same doubles
        # MultTranspose = the action of the transposed pair
        x = Q.t(np.random.default_rng(3).uniform(-1, 1, A.shape[0]))
        y = Q.capi.csr_spmv(patT, valsT, x)
        assert rel_l2(y.cpu().numpy(), A.T @ x.cpu().numpy()) < 1e-14


def test_adjoint_solve_and_shape_sensitivity(P):
    """reverseAdjointTimestep (quasi-static branch) + computeTimestepShapeSensitivity on the patch problem"""
    import scipy.sparse.linalg as spla
    prob, Q = P
    capi = Q.capi
    g = load_golden("krylov_solid_neohookean_hex_p2_patch")
    ess = np.asarray(g["ess"], dtype=np.int32)
    n = len(prob["U"])
    Q.F(0.0, [Q.S, Q.U, Q.A], wrt=1)
    pat, vals = Q.F.assemble(1)
    J = _scipy(pat, vals, n)
    patT = pat.transpose()
    valsT = patT.transpose_values(vals)
    # adjoint load (a quantity of interest's dq/du in the reference); essential entries carry the adjoint's boundary value 0
    load = np.random.default_rng(7).uniform(-1, 1, n)
    load[ess] = 0.0
    capi.eliminate_rows_cols(patT, valsT, ess, want_Ae=False)  # bc.apply(*J_T, load, adjoint_essential)
    dinv = capi.jacobi_setup(capi.csr_diagonal(patT, valsT))
    cg = capi.CGSolver(n, rel_tol=1e-12, max_iter=2000)
    lam, info = cg.Mult(patT, Q.t(load), values=valsT, dinv=dinv)
    assert info["converged"]
    # direct solve of the same eliminated system
    JT = J.T.tolil()
    for b in ess:
        JT[b, :] = 0.0
        JT[:, b] = 0.0
        JT[b, b] = 1.0
    lam_ref = spla.spsolve(JT.tocsc(), load)
    assert rel_l2(lam.cpu().numpy(), lam_ref) < 1e-8
    # shape sensitivity: (dr/dshape)^T lambda through the transposed rectangular matrix ...
    Q.F(0.0, [Q.S, Q.U, Q.A], wrt=0)
    patS, valsS = Q.F.assemble(0)
    patST = patS.transpose()
    sens = capi.csr_spmv(patST, patST.transpose_values(valsS), lam)
    assert sens.numel() == len(prob["S"])
    # ... obeys the adjoint identity with the matrix-free action dr/dshape dS
    dS = Q.t(np.random.default_rng(11).uniform(-1, 1, len(prob["S"])))
    lhs = float((lam * Q.F.GradientMult(0, dS)).sum())
    rhs = float((sens * dS).sum())
    assert abs(lhs - rhs) <= 1e-12 * max(abs(lhs), 1e-300)
