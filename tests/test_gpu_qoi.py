"""-m gpu: quantities of interest, `Functional<double(trials...)>` (src/serac/numerics/functional/functional_qoi.inl): the
test space is QOI -- one total per element (F/detail/qoi.inl), summed by the scatter -- against the reference's own QOI
kernels: per-element totals, the scalar, its directional derivative and its gradient vector (assemble(grad) of a QOI).
Tolerance 1e-12 relative."""
import glob
import os

import numpy as np
import pytest

from harness import GOLDEN, KIND_TO_NAME, load_golden, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-12
NAMES = sorted(os.path.splitext(os.path.basename(f))[0] for f in glob.glob(os.path.join(GOLDEN, "qoi_*.npz")))


@pytest.fixture(scope="module", params=NAMES)
def case(request, cuda):
    import torch
    from serac_b200 import capi
    prob = load_golden(request.param)
    geom, p, dim, ne = int(prob["geom"]), int(prob["p"]), int(prob["dim"]), int(prob["ne"])
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=cuda)
    G = capi.ElementRestriction(capi.h1_dof_info(geom, p, prob["native_dofs"]), dim, int(prob["num_nodes"]))
    # QoIElementRestriction: every element's single "dof" is entry 0 of a length-1 L-vector
    G_qoi = capi.ElementRestriction(capi.h1_dof_info(geom, 0, np.zeros((ne, 1), dtype=np.int32)), 1, 1)
    G_mesh = capi.ElementRestriction(prob["mesh_dof_info"], dim, int(prob["mesh_num_nodes"]), int(prob["mesh_ordering"]))
    nodes = t(prob["mesh_nodes"])
    X_q, J_q = capi.geometric_factors(geom, int(prob["mesh_order"]), p + 1, G_mesh, nodes)
    I = capi.Integral(f"{prob['kind']}_{KIND_TO_NAME[geom]}_p{p}", prob["params"], X_q, J_q, ne)
    F = capi.Functional(G_qoi, [G])
    F.AddDomainIntegral(I, [0])
    return prob, dict(G=G, G_qoi=G_qoi, I=I, F=F, U=t(prob["U"]), dU=t(prob["dU"]), keep=(nodes, X_q, J_q), torch=torch,
                      dev=cuda)


def test_element_totals_and_value(case):
    prob, P = case
    q_E = P["torch"].zeros(int(prob["ne"]), dtype=P["torch"].float64, device=P["dev"])
    P["I"].Mult(0.0, [P["G"].field(P["U"])], q_E)
    assert rel_l2(q_E.cpu().numpy(), prob["q_E"]) < TOL
    value = P["F"](0.0, [P["U"]])
    assert value.numel() == 1 and abs(value.item() - prob["value"]) < TOL * abs(prob["value"])


def test_directional_derivative_and_gradient_vector(case):
    prob, P = case
    value = P["F"](0.0, [P["U"]], wrt=0)
    assert abs(value.item() - prob["value"]) < TOL * abs(prob["value"])
    dq = P["F"].GradientMult(0, P["dU"])
    assert abs(dq.item() - prob["dvalue"]) < TOL * abs(prob["dvalue"])
    pat, vals = P["F"].assemble(0)  # 1 x n matrix = the gradient vector
    rp, ci = pat.to_host()
    assert rp.tolist() == [0, len(ci)]
    grad = np.zeros(P["G"].LSize())
    grad[ci] = vals.cpu().numpy()
    assert rel_l2(grad, prob["gradient"]) < TOL
    assert abs(grad @ prob["dU"] - prob["dvalue"]) < 1e-11 * abs(prob["dvalue"])


BNAMES = sorted(os.path.splitext(os.path.basename(f))[0] for f in glob.glob(os.path.join(GOLDEN, "bqoi_*.npz")))


@pytest.mark.parametrize("name", BNAMES)
def test_boundary_qoi(name, cuda):
    """Functional<double(H1<p,dim>)>::AddBoundaryIntegral (functional_qoi.inl:196-215): per-face totals, the scalar, its
    directional derivative and its gradient vector against the reference's boundary QOI kernels"""
    import torch
    from serac_b200 import capi
    prob = load_golden(name)
    geom, fg, p, dim, ne, nf = (int(prob[k]) for k in ("geom", "face_geom", "p", "dim", "ne", "nf"))
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=cuda)
    G = capi.ElementRestriction(capi.h1_dof_info(geom, p, prob["native_dofs"]), dim, int(prob["num_nodes"]))
    Gf = capi.ElementRestriction(capi.h1_dof_info(fg, p, prob["face_native_dofs"]), dim, int(prob["num_nodes"]))
    assert np.array_equal(capi.h1_dof_info(fg, p, prob["face_native_dofs"]), prob["face_dof_info"])
    G_qoi = capi.ElementRestriction(capi.h1_dof_info(geom, 0, np.zeros((ne, 1), dtype=np.int32)), 1, 1)
    Gf_qoi = capi.ElementRestriction(capi.h1_dof_info(fg, 0, np.zeros((nf, 1), dtype=np.int32)), 1, 1)
    Gf_mesh = capi.ElementRestriction(capi.h1_dof_info(fg, int(prob["mesh_order"]), prob["mesh_face_native_dofs"]), dim,
                                      int(prob["mesh_num_nodes"]), int(prob["mesh_ordering"]))
    nodes = t(prob["mesh_nodes"])
    X_q, J_q = capi.geometric_factors(fg, int(prob["mesh_order"]), p + 1, Gf_mesh, nodes)
    assert rel_l2(X_q.cpu().numpy(), prob["X_q"]) < 1e-14 and rel_l2(J_q.cpu().numpy(), prob["J_q"]) < 1e-14
    I = capi.Integral(f"{prob['kind']}_{KIND_TO_NAME[geom]}_p{p}", prob["params"], X_q, J_q, nf)
    F = capi.Functional(G_qoi, [G])
    F.AddBoundaryIntegral(I, [0], Gf_qoi, [Gf])
    U, dU = t(prob["U"]), t(prob["dU"])
    q_E = torch.zeros(nf, dtype=torch.float64, device=cuda)
    I.Mult(0.0, [Gf.field(U)], q_E)
    assert rel_l2(q_E.cpu().numpy(), prob["q_E"]) < TOL
    value = F(0.0, [U], wrt=0)
    assert abs(value.item() - prob["value"]) < TOL * abs(prob["value"])
    dq = F.GradientMult(0, dU)
    assert abs(dq.item() - prob["dvalue"]) < TOL * abs(prob["dvalue"])
    pat, vals = F.assemble(0)
    rp, ci = pat.to_host()
    grad = np.zeros(G.LSize())
    grad[ci] = vals.cpu().numpy()
    assert rel_l2(grad, prob["gradient"]) < TOL
