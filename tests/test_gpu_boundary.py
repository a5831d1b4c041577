"""-m gpu: boundary integrals (AddBoundaryIntegral: traction / follower pressure on the faces of a hex mesh) against the
reference's MakeBoundaryIntegral kernels (boundary_integral_kernels.hpp:159-325, fixtures from oracle/make_golden.py).

Bit-exact: face restriction tables.  <= 1e-12 relative L2: face geometric factors, residual, Jacobian actions, face element
matrices and their contribution to the CSR values of the domain sparsity pattern."""
import numpy as np
import pytest

from harness import BoundaryProduct, boundary_golden_names, load_golden, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module", params=boundary_golden_names())
def case(request, cuda):
    prob = load_golden(request.param)
    return prob, BoundaryProduct(prob, cuda)


def test_face_restriction_tables_bit_exact(case):
    prob, P = case
    assert np.array_equal(P.face_dof_info, prob["face_dof_info"])
    assert np.array_equal(P.shape_face_dof_info, prob["shape_face_dof_info"])


def test_face_geometric_factors(case):
    prob, P = case
    assert P.X_q.shape == prob["X_q"].shape and P.J_q.shape == prob["J_q"].shape
    assert rel_l2(P.X_q.cpu().numpy(), prob["X_q"]) < 1e-14
    assert rel_l2(P.J_q.cpu().numpy(), prob["J_q"]) < 1e-14


def test_boundary_residual(case):
    prob, P = case
    r_E = P.torch.zeros(P.Gf.ESize(), dtype=P.torch.float64, device=P.dev)
    P.integral.Mult(0.0, P.fields(), r_E)
    assert rel_l2(r_E.cpu().numpy(), prob["r_E"]) < TOL
    r = P.F(0.0, [P.S, P.U, P.A])
    assert rel_l2(r.cpu().numpy(), prob["r"]) < TOL


def test_boundary_jvp(case):
    prob, P = case
    r = P.F(0.0, [P.S, P.U, P.A], wrt=1)
    assert rel_l2(r.cpu().numpy(), prob["r"]) < TOL
    dr = P.F.GradientMult(1, P.dU).cpu().numpy()
    if np.linalg.norm(prob["dr"]) == 0.0:  # dead load: d/du is structurally zero
        assert P.integral.derivative_mask(1) == 0 and not dr.any()
    else:
        assert P.integral.derivative_mask(1) == 2  # d01 only: the normal depends on du/dxi, not on u itself
        assert rel_l2(dr, prob["dr"]) < TOL
    P.F(0.0, [P.S, P.U, P.A], wrt=0)
    dr = P.F.GradientMult(0, P.dS).cpu().numpy()
    assert rel_l2(dr, prob["dr_dshape"]) < TOL


def test_boundary_element_gradient_and_assembly(case):
    prob, P = case
    P.F(0.0, [P.S, P.U, P.A], wrt=1)
    K_e = P.torch.zeros(prob["K_e"].shape, dtype=P.torch.float64, device=P.dev)
    P.integral.ComputeElementGradients(1, K_e)
    scale = max(np.linalg.norm(prob["K_e"]), 1e-300)
    assert np.linalg.norm(K_e.cpu().numpy() - prob["K_e"]) <= TOL * scale
    pat, vals = P.F.assemble(1)
    rp, ci = pat.to_host()
    assert np.array_equal(rp, prob["row_ptr"]) and np.array_equal(ci, prob["col_ind"])
    assert np.linalg.norm(vals.cpu().numpy() - prob["values"]) <= TOL * max(np.linalg.norm(prob["values"]), 1e-300)
    if np.linalg.norm(prob["dr"]) > 0:
        import scipy.sparse as sp
        A = sp.csr_matrix((vals.cpu().numpy(), ci, rp), shape=(len(rp) - 1, len(prob["U"])))
        assert rel_l2(A @ prob["dU"], prob["dr"]) < TOL


def test_domain_plus_boundary_functional(cuda):
    """SolidMechanics' residual: one domain integral (material) + one boundary integral (pressure) in the same Functional;
    evaluate / JVP / assemble sum both contributions (functional.hpp:472-486, 418-426, 610-644)."""
    from harness import Product
    bprob = load_golden("bdr_solid_pressure_hex_p2_cart")
    B = BoundaryProduct(bprob, cuda)
    # a domain fixture on the same mesh (cart3d 2x1x2, p = 2) shares spaces and input layout
    dprob = load_golden("solid_linear_isotropic_hex_p2_cart")
    D = Product(dprob, cuda)
    assert np.array_equal(dprob["native_dofs"], bprob["native_dofs"])
    F = D.capi.Functional(D.G, [D.G_shape, D.G, D.G])
    F.AddDomainIntegral(D.integral, [0, 1, 2])
    F.AddBoundaryIntegral(B.integral, [0, 1, 2], B.Gf, [B.Gf_shape, B.Gf, B.Gf])
    S, U, A, dU = D.S, D.U, D.A, D.dU
    r = F(0.0, [S, U, A], wrt=1).cpu().numpy()
    r_sum = D.F(0.0, [S, U, A], wrt=1).cpu().numpy() + B.F(0.0, [S, U, A], wrt=1).cpu().numpy()
    assert rel_l2(r, r_sum) < 1e-14
    dr = F.GradientMult(1, dU).cpu().numpy()
    dr_sum = D.F.GradientMult(1, dU).cpu().numpy() + B.F.GradientMult(1, dU).cpu().numpy()
    assert rel_l2(dr, dr_sum) < 1e-14
    pat, vals = F.assemble(1)
    _, vd = D.F.assemble(1)
    _, vb = B.F.assemble(1)
    assert rel_l2(vals.cpu().numpy(), (vd + vb).cpu().numpy()) < 1e-14
    import scipy.sparse as sp
    rp, ci = pat.to_host()
    M = sp.csr_matrix((vals.cpu().numpy(), ci, rp), shape=(len(rp) - 1, len(rp) - 1))
    assert rel_l2(M @ dU.cpu().numpy(), dr) < TOL
