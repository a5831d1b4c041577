"""-m gpu: a discontinuous (L2) parameter field, `H1<p>(H1<p> u, L2<0> k)` -- how the physics modules take material
parameters.  The L2 restriction table is mfem's element dof list, unpermuted (element_restriction.cpp:262-268); order 0
has one node per element.  Residual, Jacobian actions and assembled matrices (square H1 x H1 and rectangular H1 x L2)
against the reference's kernels; tolerance 1e-12, CSR patterns bit-exact."""
import glob
import os

import numpy as np
import pytest

from harness import GOLDEN, KIND_TO_NAME, load_golden, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-12
NAMES = sorted(os.path.splitext(os.path.basename(f))[0] for f in glob.glob(os.path.join(GOLDEN, "l2param_*.npz")))


@pytest.mark.parametrize("name", NAMES)
def test_l2_parameter_field(name, cuda):
    import torch
    import scipy.sparse as sp
    from serac_b200 import capi
    prob = load_golden(name)
    geom, p, dim, ne = (int(prob[k]) for k in ("geom", "p", "dim", "ne"))
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=cuda)
    G = capi.ElementRestriction(capi.h1_dof_info(geom, p, prob["native_dofs"]), 1, int(prob["num_nodes"]))
    l2_info = np.ascontiguousarray(prob["l2_native_dofs"]).astype(np.uint64)  # DG dofs: no permutation
    assert np.array_equal(l2_info, prob["l2_dof_info"])
    G_l2 = capi.ElementRestriction(l2_info, 1, int(prob["l2_num_nodes"]))
    G_mesh = capi.ElementRestriction(prob["mesh_dof_info"], dim, int(prob["mesh_num_nodes"]), int(prob["mesh_ordering"]))
    nodes = t(prob["mesh_nodes"])
    X_q, J_q = capi.geometric_factors(geom, int(prob["mesh_order"]), p + 1, G_mesh, nodes)
    I = capi.Integral(f"param_diffusion_{KIND_TO_NAME[geom]}_p{p}", prob["params"], X_q, J_q, ne)
    F = capi.Functional(G, [G, G_l2])
    F.AddDomainIntegral(I, [0, 1])
    U, K = t(prob["U"]), t(prob["K"])
    r = F(0.0, [U, K])
    assert rel_l2(r.cpu().numpy(), prob["r"]) < TOL
    for wrt, d in ((0, prob["dU"]), (1, prob["dK"])):
        r = F(0.0, [U, K], wrt=wrt)
        assert rel_l2(r.cpu().numpy(), prob["r"]) < TOL
        dr = F.GradientMult(wrt, t(d)).cpu().numpy()
        assert rel_l2(dr, prob[f"dr{wrt}"]) < TOL
        pat, vals = F.assemble(wrt)
        rp, ci = pat.to_host()
        assert np.array_equal(rp, prob[f"row_ptr{wrt}"]) and np.array_equal(ci, prob[f"col_ind{wrt}"])
        assert rel_l2(vals.cpu().numpy(), prob[f"values{wrt}"]) < TOL
        A = sp.csr_matrix((vals.cpu().numpy(), ci, rp), shape=(G.LSize(), len(d)))
        assert rel_l2(A @ d, dr) < 1e-12
