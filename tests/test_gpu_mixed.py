"""-m gpu: meshes of mixed element geometry.  The reference keeps one ElementRestriction per geometry
(BlockElementRestriction, F/element_restriction.cpp:471-536), one kernel per geometry (F/integral.hpp:254-257) and builds
the sparsity pattern over all of them (F/dof_numbering.hpp:290-351).  Here: sb_csr_create_blocks against the union of the
element couplings computed with numpy (bit-exact), and a Functional with one integral per geometry
(sb_functional_add_integral_on) against the sum of the two single-geometry Functionals."""
import numpy as np
import pytest

from harness import Product, load_golden, rel_l2

pytestmark = pytest.mark.gpu


def _expected_pattern(blocks, n_rows_nodes, ct, ck):
    """rows = test vdofs (byVDIM), sorted unique trial vdofs of every element that contains the row's node"""
    rows = [set() for _ in range(n_rows_nodes)]
    for test_tab, trial_tab in blocks:
        for te, ke in zip(test_tab, trial_tab):
            for n in te:
                rows[int(n)].update(int(k) for k in ke)
    rp, ci = [0], []
    for n in range(n_rows_nodes):
        cols = sorted(rows[n])
        for _ in range(ct):
            ci.extend(k * ck + b for k in cols for b in range(ck))
            rp.append(len(ci))
    return np.asarray(rp, dtype=np.int32), np.asarray(ci, dtype=np.int32)


def test_pattern_over_blocks_is_the_union(cuda):
    from serac_b200 import capi
    # 9 nodes: two quadrilaterals below, four triangles above (conforming), vector-valued rows, scalar columns
    quads = np.array([[0, 1, 3, 4], [1, 2, 4, 5]], dtype=np.uint64)          # lexicographic quad nodes
    tris = np.array([[3, 4, 7], [3, 7, 6], [4, 5, 8], [4, 8, 7]], dtype=np.uint64)
    ct, ck = 2, 1
    Gq_t, Gt_t = capi.ElementRestriction(quads, ct, 9), capi.ElementRestriction(tris, ct, 9)
    Gq_k, Gt_k = capi.ElementRestriction(quads, ck, 9), capi.ElementRestriction(tris, ck, 9)
    pat = capi.CsrPattern([Gq_t, Gt_t], [Gq_k, Gt_k])
    rp, ci = pat.to_host()
    erp, eci = _expected_pattern([(quads, quads), (tris, tris)], 9, ct, ck)
    assert np.array_equal(rp, erp) and np.array_equal(ci, eci)
    # one block alone is a strict subset
    alone = capi.CsrPattern(Gq_t, Gq_k)
    assert alone.nnz < pat.nnz


def test_one_integral_per_block_sums_like_two_functionals(cuda):
    """two element blocks of ONE space pair (here: the even and the odd elements of a hex mesh, standing in for two
    geometries): residual, Jacobian action and assembled matrix of the blocked Functional equal those of the plain one"""
    import torch
    from serac_b200 import capi
    prob = load_golden("solid_linear_isotropic_hex_p2_cart")
    P = Product(prob, cuda)
    ne = int(prob["ne"])
    dim = int(prob["dim"])
    even, odd = np.arange(0, ne, 2), np.arange(1, ne, 2)
    blocks = []
    for sel in (even, odd):
        G = capi.ElementRestriction(P.dof_info[sel], dim, int(prob["num_nodes"]))
        Gs = capi.ElementRestriction(P.shape_dof_info[sel], dim, int(prob["shape_num_nodes"]))
        Gm = capi.ElementRestriction(np.asarray(prob["mesh_dof_info"])[sel], dim, int(prob["mesh_num_nodes"]),
                                     int(prob["mesh_ordering"]))
        X_q, J_q = capi.geometric_factors(int(prob["geom"]), int(prob["mesh_order"]), P.p + 1, Gm, P.nodes)
        I = capi.Integral(f"{prob['kind']}_hex_p{P.p}", prob["params"], X_q, J_q, len(sel))
        blocks.append((G, Gs, I, (X_q, J_q)))
    F = capi.Functional(P.G, [P.G_shape, P.G, P.G])
    for G, Gs, I, _ in blocks:
        F.AddDomainIntegralOn(I, [0, 1, 2], G, [Gs, G, G])
    r = F(0.0, [P.S, P.U, P.A], wrt=1).cpu().numpy()
    assert rel_l2(r, prob["r"]) < 1e-12
    dr = F.GradientMult(1, P.dU).cpu().numpy()
    assert rel_l2(dr, prob["dr"]) < 1e-12
    pat, vals = F.assemble(1)
    rp, ci = pat.to_host()
    assert np.array_equal(rp, prob["row_ptr"]) and np.array_equal(ci, prob["col_ind"])
    assert rel_l2(vals.cpu().numpy(), prob["values"]) < 1e-12
    torch.cuda.synchronize()
