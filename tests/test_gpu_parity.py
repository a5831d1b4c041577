"""-m gpu: parity of the CUDA path (through the C ABI) against the reference's outputs.

Tolerances: bit-exact for restriction tables, gathered E-vectors, scatter sums and the CSR pattern; <= 1e-12 relative L2
for FP64 residuals, Jacobian actions, element matrices and assembled values (BASELINE.json north_star)."""
import numpy as np
import pytest

from harness import Product, golden_names, load_golden, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module", params=golden_names())
def case(request, cuda):
    prob = load_golden(request.param)
    return prob, Product(prob, cuda)


def test_restriction_tables_bit_exact(case):
    prob, P = case
    assert np.array_equal(P.dof_info, prob["dof_info"])
    assert np.array_equal(P.shape_dof_info, prob["shape_dof_info"])


def test_gather_bit_exact(case):
    prob, P = case
    E = P.G.Gather(P.U).cpu().numpy()
    assert np.array_equal(E, prob["U_E"])


def test_scatter_add_bit_exact(case):
    prob, P = case
    L = P.torch.zeros(P.G.LSize(), dtype=P.torch.float64, device=P.dev)
    P.G.ScatterAdd(P.t(prob["r_E"]), L)
    assert np.array_equal(L.cpu().numpy(), prob["r"])
    L2 = P.torch.full((P.G.LSize(),), 7.0, dtype=P.torch.float64, device=P.dev)
    P.G.Scatter(P.t(prob["r_E"]), L2)
    assert np.array_equal(L2.cpu().numpy(), prob["r"])


def test_geometric_factors(case):
    prob, P = case
    assert rel_l2(P.X_q.cpu().numpy(), prob["X_q"]) < 1e-14
    assert rel_l2(P.J_q.cpu().numpy(), prob["J_q"]) < 1e-14


def test_residual(case):
    prob, P = case
    r_E = P.torch.zeros(P.G.ESize(), dtype=P.torch.float64, device=P.dev)
    P.integral.Mult(0.0, P.fields(), r_E)
    assert rel_l2(r_E.cpu().numpy(), prob["r_E"]) < TOL
    r = P.F(0.0, [P.S, P.U, P.A])
    assert rel_l2(r.cpu().numpy(), prob["r"]) < TOL


def test_residual_with_derivative_store_and_jvp(case):
    prob, P = case
    r_E = P.torch.zeros(P.G.ESize(), dtype=P.torch.float64, device=P.dev)
    P.integral.Mult(0.0, P.fields(), r_E, wrt=1)
    assert rel_l2(r_E.cpu().numpy(), prob["r_E_ad"]) < TOL
    dr_E = P.torch.zeros(P.G.ESize(), dtype=P.torch.float64, device=P.dev)
    P.integral.GradientMult(1, P.G.field(P.dU), dr_E)
    assert rel_l2(dr_E.cpu().numpy(), prob["dr_E"]) < TOL
    # Functional level: (r, drdu) = F(t, shape, differentiate_wrt(u), a); drdu.Mult(du)
    r = P.F(0.0, [P.S, P.U, P.A], wrt=1)
    assert rel_l2(r.cpu().numpy(), prob["r"]) < TOL
    dr = P.F.GradientMult(1, P.dU)
    assert rel_l2(dr.cpu().numpy(), prob["dr"]) < TOL


def test_jvp_wrt_acceleration_and_shape(case):
    prob, P = case
    P.F(0.0, [P.S, P.U, P.A], wrt=2)
    dr = P.F.GradientMult(2, P.dU)
    assert rel_l2(dr.cpu().numpy(), prob["dr_da"]) < TOL
    P.F(0.0, [P.S, P.U, P.A], wrt=0)
    dr = P.F.GradientMult(0, P.t(prob["dS"]))
    assert rel_l2(dr.cpu().numpy(), prob["dr_dshape"]) < TOL


def test_host_buffer_path_matches_device_path(case):
    prob, P = case
    r = P.F(0.0, [prob["S"], prob["U"], prob["A"]], wrt=1)
    assert isinstance(r, np.ndarray) and rel_l2(r, prob["r"]) < TOL
    dr = P.F.GradientMult(1, prob["dU"])
    assert rel_l2(dr, prob["dr"]) < TOL


def test_element_gradient_and_assembly(case):
    prob, P = case
    if "K_e" not in prob:
        pytest.skip("fixture without matrices")
    P.F(0.0, [P.S, P.U, P.A], wrt=1)
    K_e = P.torch.zeros(prob["K_e"].shape, dtype=P.torch.float64, device=P.dev)
    P.integral.ComputeElementGradients(1, K_e)
    assert rel_l2(K_e.cpu().numpy(), prob["K_e"]) < TOL
    pat, vals = P.F.assemble(1)
    rp, ci = pat.to_host()
    assert np.array_equal(rp, prob["row_ptr"])  # bit-exact sparsity pattern
    assert np.array_equal(ci, prob["col_ind"])
    assert rel_l2(vals.cpu().numpy(), prob["values"]) < TOL
    # check_gradient.hpp:34-44: matrix-free JVP == assembled SpMV
    import scipy.sparse as sp
    A = sp.csr_matrix((vals.cpu().numpy(), ci, rp), shape=(len(rp) - 1, len(prob["U"])))
    assert rel_l2(A @ prob["dU"], prob["dr"]) < TOL
