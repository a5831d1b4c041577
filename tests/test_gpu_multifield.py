"""-m gpu: parity of the parameterized physics signatures (thermomechanics: SolidMechanics with a temperature parameter
field and HeatTransfer with a displacement parameter field, one thermoelastic material evaluated by both -- BASELINE.json
configs[3]) against the reference's own kernels: residual, the Jacobian action and the assembled (square or rectangular)
matrix with respect to EVERY argument, and the committed quadrature-point state.  Tolerance 1e-12 relative L2; CSR
patterns bit-exact."""
import numpy as np
import pytest

from harness import MultiFieldProduct, load_golden, multifield_golden_names, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module", params=multifield_golden_names())
def case(request, cuda):
    prob = load_golden(request.param)
    return prob, MultiFieldProduct(prob, cuda)


def test_tables_and_geometry(case):
    prob, P = case
    assert np.array_equal(P.dof_info, prob["dof_info"])
    assert rel_l2(P.X_q.cpu().numpy(), prob["X_q"]) < 1e-14
    assert rel_l2(P.J_q.cpu().numpy(), prob["J_q"]) < 1e-14


def test_residual(case):
    prob, P = case
    r = P.F(0.0, P.fields)
    assert rel_l2(r.cpu().numpy(), prob["r"]) < TOL


@pytest.mark.parametrize("wrt", [0, 1, 2, 3])
def test_jacobian_action_and_matrix_wrt_every_argument(case, wrt):
    prob, P = case
    r = P.F(0.0, P.fields, wrt=wrt)
    assert rel_l2(r.cpu().numpy(), prob["r"]) < TOL
    dr = P.F.GradientMult(wrt, P.dfields[wrt])
    want = prob[f"dr{wrt}"]
    scale = max(np.linalg.norm(want), 1e-300)
    if np.linalg.norm(want) == 0.0:  # structurally present but numerically zero block (heat flux vs displacement)
        assert np.abs(dr.cpu().numpy()).max() == 0.0
    else:
        assert np.linalg.norm(dr.cpu().numpy() - want) / scale < TOL
    pat, vals = P.F.assemble(wrt)
    rp, ci = pat.to_host()
    assert np.array_equal(rp, prob[f"row_ptr{wrt}"]) and np.array_equal(ci, prob[f"col_ind{wrt}"])
    v = vals.cpu().numpy()
    if np.linalg.norm(prob[f"values{wrt}"]) == 0.0:
        assert np.abs(v).max() == 0.0
    else:
        assert rel_l2(v, prob[f"values{wrt}"]) < TOL
    # the assembled matrix applied to du equals the matrix-free action (check_gradient.hpp:16-50)
    import scipy.sparse as sp
    A = sp.csr_matrix((v, ci, rp), shape=(len(rp) - 1, P.G_trials[wrt].LSize()))
    assert np.linalg.norm(A @ prob[f"dfield{wrt}"] - dr.cpu().numpy()) <= 1e-11 * max(scale, 1.0)


def test_state_commit(case):
    prob, P = case
    if "qdata1" not in prob:
        pytest.skip("stateless integrand")
    P.qdata.copy_(P.t(prob["qdata0"]))
    P.F(0.0, P.fields, update_state=True)
    got = P.qdata.cpu().numpy().ravel()
    want = np.asarray(prob["qdata1"]).ravel().view(np.float64) if prob["qdata1"].dtype != np.float64 else prob["qdata1"].ravel()
    assert rel_l2(got, want) < TOL
    r_E = P.torch.zeros(P.G_test.ESize(), dtype=P.torch.float64, device=P.dev)
    P.integral.Mult(0.0, [G.field(f) for G, f in zip(P.G_trials, P.fields)], r_E)
    assert rel_l2(r_E.cpu().numpy(), prob["r_E_after_commit"]) < TOL
    P.qdata.copy_(P.t(prob["qdata0"]))
