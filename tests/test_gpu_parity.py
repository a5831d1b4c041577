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
    sel = slice(None) if P.domain is None else P.domain  # the fixture holds the factors of every element
    assert rel_l2(P.X_q.cpu().numpy(), prob["X_q"][sel]) < 1e-14
    assert rel_l2(P.J_q.cpu().numpy(), prob["J_q"][sel]) < 1e-14


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


def test_state_commit(case):
    """QuadratureData semantics (domain_integral_kernels.hpp:143-147): state is read every evaluation, written back only
    when update_state is set; the next evaluation then starts from the committed state."""
    prob, P = case
    if "qdata1" not in prob:
        pytest.skip("stateless integrand")
    r_E = P.torch.zeros(P.G.ESize(), dtype=P.torch.float64, device=P.dev)
    P.integral.Mult(0.0, P.fields(), r_E)  # no commit: state untouched
    assert np.array_equal(P.qdata.cpu().numpy().ravel(), np.asarray(prob["qdata0"]).ravel())
    P.integral.Mult(0.0, P.fields(), r_E, update_state=True)
    assert rel_l2(r_E.cpu().numpy(), prob["r_E"]) < TOL
    assert rel_l2(P.qdata.cpu().numpy(), prob["qdata1"]) < TOL
    P.integral.Mult(0.0, P.fields(), r_E)
    assert rel_l2(r_E.cpu().numpy(), prob["r_E_after_commit"]) < TOL
    # committing through the derivative-storing kernel gives the same state
    P.qdata.copy_(P.t(prob["qdata0"]).reshape(P.qdata.shape))
    P.integral.Mult(0.0, P.fields(), r_E, wrt=1, update_state=True)
    assert rel_l2(P.qdata.cpu().numpy(), prob["qdata1"]) < TOL
    P.qdata.copy_(P.t(prob["qdata0"]).reshape(P.qdata.shape))


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


def test_tangent_fast_path_matches_dual_number_path(cuda):
    """The closed-form flux tangent (MaterialStressFunctor::flux_tangent) against the same integrand differentiated with
    dual numbers end to end (the reference's scheme, registered as *_ad): stored derivatives, residual, JVP."""
    prob = load_golden("solid_neohookean_hex_p2_patch")
    fast, ad = Product(prob, cuda), Product(prob, cuda, name_suffix="_ad")
    assert fast.integral.derivative_mask(1) & 15 == ad.integral.derivative_mask(1) == 8
    r_f = fast.F(0.0, [fast.S, fast.U, fast.A], wrt=1).cpu().numpy()
    r_a = ad.F(0.0, [ad.S, ad.U, ad.A], wrt=1).cpu().numpy()
    assert rel_l2(r_f, r_a) < 1e-14 and rel_l2(r_a, prob["r"]) < TOL
    D_f = fast.integral.export_derivatives(1).cpu().numpy()
    D_a = ad.integral.export_derivatives(1).cpu().numpy()
    assert rel_l2(D_f, D_a) < 1e-13
    assert rel_l2(ad.F.GradientMult(1, ad.dU).cpu().numpy(), prob["dr"]) < TOL
    assert rel_l2(fast.F.GradientMult(1, fast.dU).cpu().numpy(), prob["dr"]) < TOL


def test_config_a_full_size(cuda):
    """BASELINE.json configs[1] at full size (32 768 hexes, 839 619 vdofs): sampled reference residual / Jacobian action,
    plus size-independent properties: linearity of the Jacobian action and JVP == directional finite difference."""
    prob = load_golden("config_a_beam_hex_r4_p2")
    P = Product(prob, cuda)
    stride = int(prob["sample_stride"])
    r = P.F(0.0, [P.S, P.U, P.A], wrt=1).cpu().numpy()
    dr = P.F.GradientMult(1, P.dU).cpu().numpy()
    assert rel_l2(r[::stride], prob["r_sample"]) < TOL
    assert rel_l2(dr[::stride], prob["dr_sample"]) < TOL
    assert abs(np.linalg.norm(r) - prob["r_norm"]) < TOL * prob["r_norm"]
    assert abs(np.linalg.norm(dr) - prob["dr_norm"]) < TOL * prob["dr_norm"]
    # linearity: J(a x + b y) = a J x + b J y
    y = P.torch.linspace(-1, 1, P.G.LSize(), dtype=P.torch.float64, device=cuda)
    lhs = P.F.GradientMult(1, 0.3 * P.dU - 1.7 * y).cpu().numpy()
    rhs = 0.3 * dr - 1.7 * P.F.GradientMult(1, y).cpu().numpy()
    assert rel_l2(lhs, rhs) < 1e-13
    # central finite difference of the residual in direction dU (check_gradient.hpp:60-90)
    eps = 1e-6
    rp = P.F(0.0, [P.S, P.U + eps * P.dU, P.A]).cpu().numpy()
    rm = P.F(0.0, [P.S, P.U - eps * P.dU, P.A]).cpu().numpy()
    assert rel_l2((rp - rm) / (2 * eps), dr) < 1e-7


def test_config_b_properties_p3_cube(cuda):
    """BASELINE.json configs[4] (structured hex cube, H1 p=3, NeoHookean) at a size the GPU suite can afford (32^3 hexes,
    2.7 M vdofs): the size-independent properties tools/config_b.py checks at full size."""
    import sys
    sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1] / "tools"))
    import config_b
    P = config_b.build(32, cuda)
    props = config_b.check_properties(P, cuda)
    assert props["sum_of_internal_forces_rel"] < 1e-9, props  # a sum over every node: roundoff grows with the mesh
    assert props["linearity_rel"] < 1e-12, props
    assert props["symmetry_rel"] < 1e-10, props
    assert props["finite_difference_rel"] < 1e-6, props


def test_node_major_e_vector_layout_is_bit_identical(case):
    """sb_functional_* keeps its E-vector [element][node][comp]; the sums and their order are the reference's, so the
    scattered L-vector is bit-identical to the [element][comp][node] route"""
    prob, P = case
    nE, ct, nd = int(prob["ne"]), P.vdim, P.dof_info.shape[1]
    r_ref = P.torch.zeros(P.G.ESize(), dtype=P.torch.float64, device=P.dev)
    r_nm = P.torch.zeros_like(r_ref)
    P.integral.Mult(0.0, P.fields(), r_ref, wrt=1)
    P.integral.Mult(0.0, P.fields(), r_nm, wrt=1, node_major=True)
    assert P.torch.equal(r_nm.view(nE, nd, ct).permute(0, 2, 1).contiguous().view(-1), r_ref)
    L_ref = P.G.Scatter(r_ref, P.torch.empty(P.G.LSize(), dtype=P.torch.float64, device=P.dev))
    L_nm = P.G.Scatter(r_nm, P.torch.empty(P.G.LSize(), dtype=P.torch.float64, device=P.dev), node_major=True)
    assert P.torch.equal(L_ref, L_nm)
    d_ref = P.integral.GradientMult(1, P.G.field(P.dU), P.torch.zeros_like(r_ref))
    d_nm = P.integral.GradientMult(1, P.G.field(P.dU), P.torch.zeros_like(r_ref), node_major=True)
    assert P.torch.equal(d_nm.view(nE, nd, ct).permute(0, 2, 1).contiguous().view(-1), d_ref)
    acc = P.torch.full((P.G.LSize(),), 0.5, dtype=P.torch.float64, device=P.dev)
    acc2 = acc.clone()
    assert P.torch.equal(P.G.ScatterAdd(d_ref, acc), P.G.ScatterAdd(d_nm, acc2, node_major=True))


def test_async_host_outputs_match(case):
    """SB_HOST_ASYNC: results copied on the Functional's copy stream are, after wait(), bit-identical to the device path,
    also when calls are queued back to back on pinned buffers"""
    prob, P = case
    torch = P.torch
    pin = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).pin_memory()
    hS, hU, hA, hdU = (pin(prob[k]) for k in ("S", "U", "A", "dU"))
    hr, hdr = (torch.empty(P.G.LSize(), dtype=torch.float64).pin_memory() for _ in range(2))
    r = P.F(0.0, [P.S, P.U, P.A], wrt=1).cpu().numpy()
    dr = P.F.GradientMult(1, P.dU).cpu().numpy()
    for _ in range(3):  # queued back to back: the staging buffers are guarded by events
        P.F(0.0, [hS.numpy(), hU.numpy(), hA.numpy()], wrt=1, out=hr.numpy(), async_host=True)
        P.F.GradientMult(1, hdU.numpy(), out=hdr.numpy(), async_host=True)
    P.F.wait()
    assert np.array_equal(hr.numpy(), r) and np.array_equal(hdr.numpy(), dr)


def test_complementary_subdomains_sum_to_the_whole(cuda):
    """Two integrals of one Functional on complementary parts of the mesh (functional.hpp:446-498 sums every integral's
    E-vector into the same output) equal the integral on the entire mesh; an empty domain contributes exactly zero."""
    import torch
    from serac_b200 import capi
    prob = load_golden("solid_neohookean_hex_p2_subdomain")
    part = Product(prob, cuda)  # integral on the fixture's sub-domain
    whole_prob = dict(prob)
    del whole_prob["domain_elements"]
    whole = Product(whole_prob, cuda)
    ne = int(prob["ne"])
    rest = np.setdiff1d(np.arange(ne, dtype=np.int32), part.domain).astype(np.int32)
    rest_dev = torch.as_tensor(rest, dtype=torch.int32, device=cuda)
    name = "solid_neohookean_hex_p2"
    X2, J2 = capi.geometric_factors(part.geom, int(prob["mesh_order"]), 3, part.G_mesh, part.nodes, num_elements=len(rest),
                                    elements=rest_dev)
    I_rest = capi.Integral(name, prob["params"], X2, J2, len(rest), elements=rest)
    X0, J0 = capi.geometric_factors(part.geom, int(prob["mesh_order"]), 3, part.G_mesh, part.nodes, num_elements=0,
                                    elements=torch.zeros(0, dtype=torch.int32, device=cuda))
    I_empty = capi.Integral(name, prob["params"], X0, J0, 0, elements=np.zeros(0, dtype=np.int32))
    F2 = capi.Functional(part.G, [part.G_shape, part.G, part.G])
    F2.AddDomainIntegral(part.integral, [0, 1, 2])
    F2.AddDomainIntegral(I_rest, [0, 1, 2])
    F2.AddDomainIntegral(I_empty, [0, 1, 2])
    r2 = F2(0.0, [part.S, part.U, part.A], wrt=1).cpu().numpy()
    rw = whole.F(0.0, [whole.S, whole.U, whole.A], wrt=1).cpu().numpy()
    assert rel_l2(r2, rw) < 1e-14
    d2 = F2.GradientMult(1, part.dU).cpu().numpy()
    dw = whole.F.GradientMult(1, whole.dU).cpu().numpy()
    assert rel_l2(d2, dw) < 1e-14
    _, v2 = F2.assemble(1)
    _, vw = whole.F.assemble(1)
    assert rel_l2(v2.cpu().numpy(), vw.cpu().numpy()) < 1e-14
    # the empty domain alone: exactly zero residual
    F0 = capi.Functional(part.G, [part.G_shape, part.G, part.G])
    F0.AddDomainIntegral(I_empty, [0, 1, 2])
    assert float(F0(0.0, [part.S, part.U, part.A]).abs().max()) == 0.0


def test_evaluation_in_two_halves_is_bit_identical(case):
    """sb_functional_evaluate_begin / _end (elements [0, split) first, the rest later -- what a partitioned run does
    around its input exchange) returns exactly what the one-call evaluation returns, for any split point"""
    prob, P = case
    if P.domain is not None:
        pytest.skip("the split applies to whole-mesh domain integrals")
    ne = int(prob["ne"])
    ref = P.F(0.0, [P.S, P.U, P.A], wrt=1).clone()
    ref_d = P.integral.export_derivatives(1).clone()
    for split in sorted({0, 1, ne // 2, ne - 1, ne, ne + 5}):
        P.F.evaluate_begin(0.0, [P.S, P.U, P.A], wrt=1, split_element=split)
        got = P.F.evaluate_end()
        assert P.torch.equal(got, ref), f"split at {split}"
        assert P.torch.equal(P.integral.export_derivatives(1), ref_d)
    dr = P.F.GradientMult(1, P.dU).cpu().numpy()
    assert rel_l2(dr, prob["dr"]) < TOL


def test_geometric_factors_refresh_after_node_move():
    """Shape-optimisation loops move the mesh nodes and re-run GeometricFactors: the integral
    keeps POINTERS to X_q / J_q, so recomputing them in place on the device updates it -- no rebuild.  Moving the nodes by
    a field p and evaluating with zero shape displacement must equal evaluating on the original nodes with shape
    displacement p (the identity ShapeAwareFunctional rests on, shape_aware_functional.hpp:364-438)."""
    import torch
    from serac_b200 import capi
    prob = load_golden("solid_neohookean_hex_p2_patch")
    dev = torch.device("cuda:0")
    P = Product(prob, dev)
    rng = np.random.default_rng(5)
    shape = P.t(0.02 * rng.uniform(-1, 1, P.G_shape.LSize()))
    r_shape = P.F(0.0, [shape, P.U, P.A]).clone()
    # the mesh-nodes space is order 1 with the same numbering as the shape space; both byVDIM in this fixture
    assert int(prob["mesh_ordering"]) == capi.BYVDIM and int(prob["mesh_order"]) == 1
    moved = P.nodes + shape
    check = capi.check
    check(capi.lib().sb_geometric_factors(P.geom, 1, P.p + 1, P.G_mesh.h, capi._dev_ptr(moved), None, int(prob["ne"]),
                                          capi._dev_ptr(P.X_q), capi._dev_ptr(P.J_q), capi._stream_ptr()))
    r_moved = P.F(0.0, [torch.zeros_like(shape), P.U, P.A])
    assert rel_l2(r_moved.cpu().numpy(), r_shape.cpu().numpy()) < 1e-12
    assert rel_l2(r_moved.cpu().numpy(), prob["r"]) > 1e-3  # the move really changed the answer


def test_symmetric_tangent_storage(cuda):
    """hyperelastic flux tangents are stored as their upper triangle (45 of 81 doubles per point), chosen
    per integral from the functor's geometric-nonlinearity flag; everything else stays in the reference's full layout.
    The exported (expanded) derivative is exactly symmetric, and equal to the dual-number derivative to rounding
    (test_tangent_fast_path_matches_dual_number_path)."""
    prob = load_golden("solid_neohookean_hex_p2_patch")
    assert float(prob["params"][3]) == 1.0  # geometric nonlinearity on
    P = Product(prob, cuda)
    I, capi = P.integral, P.capi
    assert (I.derivative_entries(1), I.derivative_stored_entries(1), I.derivative_mask(1)) == (81, 45, 8 | 16)
    assert I.derivative_stored_entries(0) == I.derivative_entries(0)  # w.r.t. the shape field: not symmetric, full
    P.F(0.0, [P.S, P.U, P.A], wrt=1)
    D = I.export_derivatives(1).cpu().numpy().reshape(-1, 9, 9)
    assert np.abs(D).max() > 0 and np.array_equal(D, D.transpose(0, 2, 1))
    ne = int(prob["ne"])
    # NeoHookean without geometric nonlinearity: flux = Cauchy stress of du_dX, whose tangent is not symmetric -> full
    off = list(np.asarray(prob["params"], dtype=float))
    off[3] = 0.0
    I_off = capi.Integral("solid_neohookean_hex_p2", off, P.X_q, P.J_q, ne)
    assert (I_off.derivative_stored_entries(1), I_off.derivative_mask(1)) == (81, 8)
    # linear isotropic elasticity: symmetric small-strain moduli -> upper triangle without geometric nonlinearity only
    I_lin0 = capi.Integral("solid_linear_isotropic_hex_p2", off, P.X_q, P.J_q, ne)
    I_lin1 = capi.Integral("solid_linear_isotropic_hex_p2", list(prob["params"]), P.X_q, P.J_q, ne)
    assert (I_lin0.derivative_stored_entries(1), I_lin0.derivative_mask(1)) == (45, 8 | 16)
    assert (I_lin1.derivative_stored_entries(1), I_lin1.derivative_mask(1)) == (81, 8)
    # ... and the dual-number twin of the integrand never uses it
    ad = Product(prob, cuda, name_suffix="_ad")
    assert ad.integral.derivative_stored_entries(1) == 81
