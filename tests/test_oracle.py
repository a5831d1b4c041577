"""CPU: pins the numpy restatement (oracle/functional_oracle.py) against the golden fixtures that the reference's own
kernels produced (oracle/make_golden.py via oracle/_ref), and -- when oracle/_ref is present -- live against the reference
for integrands that have no fixture."""
import numpy as np
import pytest

from harness import golden_names, load_golden, rel_l2
from oracle import functional_oracle as fo
from oracle import ref

MASK = np.uint64(0x0000FFFFFFFFFFFF)


def run_oracle(prob, wrt=1, state=None):
    geom, p, dim, vdim = int(prob["geom"]), int(prob["p"]), int(prob["dim"]), int(prob["vdim"])
    q = p + 1
    table = fo.dof_info(geom, p, prob["native_dofs"])
    shape_table = fo.dof_info(geom, 1, prob["shape_native_dofs"])
    mesh_table = prob["mesh_dof_info"] & MASK
    byvdim = int(prob["mesh_ordering"]) == 1
    X_q, J_q = fo.geometric_factors(fo.gather(mesh_table, prob["mesh_nodes"], dim, byvdim), int(prob["mesh_order"]), q,
                                    dim, geom)
    qf = fo.make_qf(str(prob["kind"]), list(prob["params"]))
    trials = [fo.gather(shape_table, prob["S"], dim), fo.gather(table, prob["U"], vdim),
              fo.gather(table, prob["A"], vdim)]
    if state is None and "qdata0" in prob:
        state = j2_state(prob["qdata0"], int(prob["ne"]), fo.num_quadrature_points(geom, q), tensor_key(prob))
    r_E, blocks = fo.evaluation_kernel(qf, trials, [1, p, p], X_q, J_q, p, q, dim, wrt=wrt, state=state, geom=geom)
    outside = np.ones(int(prob["ne"]), dtype=bool)
    if "domain_elements" in prob:  # Domain::ofElements: elements outside the domain contribute nothing
        outside[prob["domain_elements"]] = False
        r_E[outside] = 0.0
        if blocks is not None:
            for b in blocks.values():
                b[outside] = 0.0
    return dict(table=table, shape_table=shape_table, X_q=X_q, J_q=J_q, r_E=r_E, blocks=blocks, p=p, q=q, dim=dim,
                vdim=vdim, geom=geom)


def tensor_key(prob):
    """name of the 3x3 state tensor: J2SmallStrain::State::plastic_strain (solid_material.hpp:212-215) or
    J2::State::Fpinv (:272-275); both are followed by the accumulated plastic strain"""
    return "Fpinv" if str(prob["kind"]) == "solid_j2_linear" else "plastic_strain"


def j2_state(qdata, ne, nq, key="plastic_strain"):
    """[ne][nq] AoS of {tensor[3][3], eqps} -> oracle dict"""
    s = np.asarray(qdata).reshape(ne, nq, 10)
    return {key: [[s[:, :, 3 * i + j].copy() for j in range(3)] for i in range(3)], "eqps": s[:, :, 9].copy()}


@pytest.fixture(scope="module", params=golden_names())
def case(request):
    prob = load_golden(request.param)
    return prob, run_oracle(prob)


def test_tables_bit_exact(case):
    prob, o = case
    assert np.array_equal(o["table"], prob["dof_info"])
    assert np.array_equal(o["shape_table"], prob["shape_dof_info"])


def test_gather_scatter_bit_exact(case):
    prob, o = case
    E = fo.gather(o["table"], prob["U"], o["vdim"])
    assert np.array_equal(E.ravel(), prob["U_E"])
    L = fo.scatter_add(o["table"], prob["r_E"], np.zeros(len(prob["U"])), o["vdim"])
    assert np.array_equal(L, prob["r"])


def test_geometric_factors(case):
    prob, o = case
    assert rel_l2(o["X_q"], prob["X_q"]) < 1e-14 and rel_l2(o["J_q"], prob["J_q"]) < 1e-14


def test_residual_and_jvp(case):
    prob, o = case
    assert rel_l2(o["r_E"], prob["r_E"]) < 1e-13
    dU_E = fo.gather(o["table"], prob["dU"], o["vdim"])
    dr_E = fo.action_of_gradient_kernel(o["blocks"], dU_E, o["p"], o["p"], o["q"], o["dim"], o["geom"])
    assert rel_l2(dr_E, prob["dr_E"]) < 1e-13


def test_state_update(case):
    """domain_integral_kernels.hpp:143-147: the state the q-function computed is what a committing evaluation stores"""
    prob, _ = case
    if "qdata1" not in prob:
        pytest.skip("stateless integrand")
    ne, nq = int(prob["ne"]), fo.num_quadrature_points(int(prob["geom"]), int(prob["p"]) + 1)
    key = tensor_key(prob)
    st = j2_state(prob["qdata0"], ne, nq, key)
    run_oracle(prob, wrt=-1, state=st)
    new = st["_updated"]
    want = np.asarray(prob["qdata1"]).reshape(ne, nq, 10)
    assert rel_l2(new["eqps"], want[:, :, 9]) < 1e-13
    for i in range(3):
        for j in range(3):
            assert np.abs(new[key][i][j] - want[:, :, 3 * i + j]).max() < 1e-13


def test_other_arguments(case):
    prob, _ = case
    o2 = run_oracle(prob, wrt=2)
    dU_E = fo.gather(o2["table"], prob["dU"], o2["vdim"])
    dr = fo.scatter_add(o2["table"], fo.action_of_gradient_kernel(o2["blocks"], dU_E, o2["p"], o2["p"], o2["q"],
                                                                 o2["dim"], o2["geom"]), np.zeros(len(prob["U"])),
                        o2["vdim"])
    assert rel_l2(dr, prob["dr_da"]) < 1e-13
    o0 = run_oracle(prob, wrt=0)
    dS_E = fo.gather(o0["shape_table"], prob["dS"], o0["dim"])
    dr = fo.scatter_add(o0["table"], fo.action_of_gradient_kernel(o0["blocks"], dS_E, 1, o0["p"], o0["q"], o0["dim"],
                                                                 o0["geom"]), np.zeros(len(prob["U"])), o0["vdim"])
    assert rel_l2(dr, prob["dr_dshape"]) < 1e-13


def test_element_matrices_and_csr(case):
    prob, o = case
    if "K_e" not in prob:
        pytest.skip("fixture without matrices")
    K_e = fo.element_gradient_kernel(o["blocks"], o["p"], o["p"], o["q"], o["dim"], o["geom"])
    assert rel_l2(K_e, prob["K_e"]) < 1e-13
    vd = fo.element_vdofs(o["table"], o["vdim"], int(prob["num_nodes"]))
    row_ptr, col_ind, keys = fo.csr_pattern(vd, vd, len(prob["U"]))
    assert np.array_equal(row_ptr, prob["row_ptr"]) and np.array_equal(col_ind, prob["col_ind"])
    vals = fo.csr_assemble(prob["K_e"], vd, vd, keys)
    assert np.array_equal(vals, prob["values"])  # same additions in the same order


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("kind,vdim,params", [("thermal_linear_conductor", 1, [1.3, 0.7, 2.1])])
def test_live_reference_other_integrands(kind, vdim, params):
    from oracle.make_golden import reference_problem
    prob = reference_problem(mesh_spec=("cart3d", 2, 2, 2, 0.03), kind=kind, p=2, params=params, vdim=vdim,
                             shape_amp=0.01)
    o = run_oracle(prob)
    assert rel_l2(o["r_E"], prob["r_E"]) < 1e-13
    dU_E = fo.gather(o["table"], prob["dU"], vdim)
    assert rel_l2(fo.action_of_gradient_kernel(o["blocks"], dU_E, 2, 2, 3, 3), prob["dr_E"]) < 1e-13
    K_e = fo.element_gradient_kernel(o["blocks"], 2, 2, 3, 3)
    assert rel_l2(K_e, prob["K_e"]) < 1e-13


# ----------------------------------------------------------------------------------------------------------------------
# quantities of interest (F/functional_qoi.inl, F/detail/qoi.inl): numpy restatement against the reference's QOI kernels
# ----------------------------------------------------------------------------------------------------------------------
def _qoi_names():
    import glob
    import os
    from harness import GOLDEN
    return sorted(os.path.splitext(os.path.basename(f))[0] for f in glob.glob(os.path.join(GOLDEN, "qoi_*.npz")))


def _quadrature_weights(geom, q, dim):
    if fo.is_simplex(geom):
        return fo.simplex_quadrature(geom, q)[1]
    w1 = fo.basis_tables(1, q)[2]
    w = w1
    for _ in range(dim - 1):
        w = np.multiply.outer(w1, w)  # Q = (qz*q + qy)*q + qx
    return w.ravel()


@pytest.mark.parametrize("name", _qoi_names())
def test_qoi_value_and_gradient(name):
    """element totals = sum_q f(x_q, u_q, grad u_q) det(J_q) w_q (qoi.inl:33-75); the gradient vector by the chain rule
    through the interpolation (functional_qoi.inl:411-455)"""
    prob = load_golden(name)
    geom, p, dim = int(prob["geom"]), int(prob["p"]), int(prob["dim"])
    q = p + 1
    a, b = (float(x) for x in prob["params"])
    table = fo.dof_info(geom, p, prob["native_dofs"])
    assert np.array_equal(table, prob["dof_info"])
    mesh_table = prob["mesh_dof_info"] & MASK
    byvdim = int(prob["mesh_ordering"]) == 1
    X_q, J_q = fo.geometric_factors(fo.gather(mesh_table, prob["mesh_nodes"], dim, byvdim), int(prob["mesh_order"]), q,
                                    dim, geom)
    assert rel_l2(X_q, prob["X_q"]) < 1e-14 and rel_l2(J_q, prob["J_q"]) < 1e-14
    U_E = fo.gather(table, prob["U"], dim)
    val, grad = fo.interpolate(U_E, p, q, dim, geom)  # [ne][c][nq], parent gradient [ne][c][dim][nq]
    Jm = np.transpose(J_q, (0, 3, 2, 1))  # [ne][nq][i][j] = dX_i/dxi_j
    invJ = np.linalg.inv(Jm)
    dv = np.linalg.det(Jm)
    du = np.einsum("ecjq,eqjk->eqck", grad, invJ)  # physical gradient
    u = np.transpose(val, (0, 2, 1))  # [ne][nq][c]
    w = _quadrature_weights(geom, q, dim)
    x0, x1 = X_q[:, 0, :], X_q[:, 1, :]
    f = a * (1.0 + x0) * np.einsum("eqc,eqc->eq", u, u) + 0.5 * b * np.einsum("eqck,eqck->eq", du, du) + np.sin(x1)
    q_E = np.einsum("eq,eq,q->e", f, dv, w)
    assert rel_l2(q_E, prob["q_E"]) < 1e-13
    assert abs(q_E.sum() - prob["value"]) < 1e-13 * abs(prob["value"])
    # d f / d u = 2 a (1 + x0) u ; d f / d (grad u) = b grad u  ->  gradient w.r.t. the dofs through B and G
    phi, dphi = fo.interpolate(np.eye(U_E.shape[2])[None], p, q, dim, geom)
    phi, dphi = phi[0], dphi[0]  # [k][nq], [k][dim][nq]
    dphi_phys = np.einsum("kjq,eqjl->ekql", dphi, invJ)
    gE = (np.einsum("eq,eqc,kq,eq,q->eck", 2.0 * a * (1.0 + x0), u, phi, dv, w) +
          b * np.einsum("eqcl,ekql,eq,q->eck", du, dphi_phys, dv, w))
    g = fo.scatter_add(table, gE, np.zeros(len(prob["U"])), dim)
    assert rel_l2(g, prob["gradient"]) < 1e-13
    assert abs(g @ prob["dU"] - prob["dvalue"]) < 1e-12 * abs(prob["dvalue"])


# ----------------------------------------------------------------------------------------------------------------------
# thermomechanics (parameterized signatures, BASELINE.json configs[3]): numpy restatement against the reference's kernels
# ----------------------------------------------------------------------------------------------------------------------
from harness import multifield_golden_names  # noqa: E402


@pytest.mark.parametrize("name", multifield_golden_names())
def test_thermomechanics_residual_jvp_and_state(name):
    prob = load_golden(name)
    geom, p, dim = int(prob["geom"]), int(prob["p"]), int(prob["dim"])
    q = p + 1
    vdims = [int(v) for v in prob["trial_vdims"]]
    test_vdim = int(prob["test_vdim"])
    table = fo.dof_info(geom, p, prob["native_dofs"])
    shape_table = fo.dof_info(geom, 1, prob["shape_native_dofs"])
    mesh_table = prob["mesh_dof_info"] & MASK
    byvdim = int(prob["mesh_ordering"]) == 1
    X_q, J_q = fo.geometric_factors(fo.gather(mesh_table, prob["mesh_nodes"], dim, byvdim), int(prob["mesh_order"]), q,
                                    dim, geom)
    tables = [shape_table] + [table] * (len(vdims) - 1)
    orders = [1] + [p] * (len(vdims) - 1)
    trials = [fo.gather(t, prob[f"field{i}"], v) for i, (t, v) in enumerate(zip(tables, vdims))]
    qf = fo.make_qf(str(prob["kind"]), list(prob["params"]))
    ne, nq = int(prob["ne"]), fo.num_quadrature_points(geom, q)

    def state0():
        if "qdata0" not in prob:
            return None
        return dict(strain_trace=np.asarray(prob["qdata0"]).reshape(ne, nq).copy())

    st = state0()
    r_E, _ = fo.evaluation_kernel(qf, trials, orders, X_q, J_q, p, q, dim, wrt=-1, state=st, geom=geom)
    assert rel_l2(r_E, prob["r_E"]) < 1e-13
    if st is not None:  # the committed state: trace of the Green strain (green_saint_venant_thermoelastic.hpp:64)
        assert rel_l2(st["_updated"]["strain_trace"], np.asarray(prob["qdata1"]).reshape(ne, nq)) < 1e-13
    for wrt in range(len(vdims)):
        _, blocks = fo.evaluation_kernel(qf, trials, orders, X_q, J_q, p, q, dim, wrt=wrt, state=state0(), geom=geom)
        du_E = fo.gather(tables[wrt], prob[f"dfield{wrt}"], vdims[wrt])
        dr_E = fo.action_of_gradient_kernel(blocks, du_E, orders[wrt], p, q, dim, geom)
        dr = fo.scatter_add(table, dr_E, np.zeros(int(prob["num_nodes"]) * test_vdim), test_vdim)
        want = prob[f"dr{wrt}"]
        if np.linalg.norm(want) == 0.0:
            assert np.abs(dr).max() < 1e-13
        else:
            assert rel_l2(dr, want) < 1e-13


# ----------------------------------------------------------------------------------------------------------------------
# boundary integrals (traction / follower pressure on quadrilateral and triangular faces)
# ----------------------------------------------------------------------------------------------------------------------
from harness import boundary_golden_names  # noqa: E402


@pytest.mark.parametrize("name", boundary_golden_names())
def test_boundary_residual_and_jvps(name):
    prob = load_golden(name)
    fg, p, dim = int(prob["face_geom"]), int(prob["p"]), int(prob["dim"])
    q = p + 1
    face_table = fo.dof_info(fg, p, prob["face_native_dofs"])
    shape_face_table = fo.dof_info(fg, 1, prob["shape_face_native_dofs"])
    assert np.array_equal(face_table, prob["face_dof_info"])
    assert np.array_equal(shape_face_table, prob["shape_face_dof_info"])
    mesh_face_table = fo.dof_info(fg, int(prob["mesh_order"]), prob["mesh_face_native_dofs"])
    byvdim = int(prob["mesh_ordering"]) == 1
    nodes_E = fo.gather(mesh_face_table, prob["mesh_nodes"], dim, byvdim)
    Xv, Xg = fo.interpolate(nodes_E, int(prob["mesh_order"]), q, dim - 1, fg)  # geometric_factors.cpp:140-220
    X_q, J_q = Xv, np.transpose(Xg, (0, 2, 1, 3))
    assert rel_l2(X_q, prob["X_q"]) < 1e-14 and rel_l2(J_q, prob["J_q"]) < 1e-14
    qf = fo.make_boundary_qf(str(prob["kind"]), list(prob["params"]))
    trials = [fo.gather(shape_face_table, prob["S"], dim), fo.gather(face_table, prob["U"], dim),
              fo.gather(face_table, prob["A"], dim)]
    orders = [1, p, p]
    r_E, _ = fo.boundary_evaluation_kernel(qf, trials, orders, X_q, J_q, p, q, fg)
    assert rel_l2(r_E.ravel(), prob["r_E"]) < 1e-13
    r = fo.scatter_add(face_table, r_E, np.zeros(len(prob["U"])), dim)
    assert rel_l2(r, prob["r"]) < 1e-13
    for wrt, table, key, dkey, order in ((1, face_table, "dr", "dU", p), (0, shape_face_table, "dr_dshape", "dS", 1)):
        _, blocks = fo.boundary_evaluation_kernel(qf, trials, orders, X_q, J_q, p, q, fg, wrt=wrt)
        d_E = fo.boundary_action_of_gradient_kernel(blocks, fo.gather(table, prob[dkey], dim), order, p, q, fg)
        dr = fo.scatter_add(face_table, d_E, np.zeros(len(prob["U"])), dim)
        if np.linalg.norm(prob[key]) == 0.0:
            assert np.abs(dr).max() < 1e-13
        else:
            assert rel_l2(dr, prob[key]) < 1e-13


def _bqoi_names():
    import glob
    import os
    from harness import GOLDEN
    return sorted(os.path.splitext(os.path.basename(f))[0] for f in glob.glob(os.path.join(GOLDEN, "bqoi_*.npz")))


@pytest.mark.parametrize("name", _bqoi_names())
def test_boundary_qoi_value(name):
    """boundary quantity of interest: sum over faces and points of f(x, u, du/dxi) |dX/dxi_1 x dX/dxi_2| w"""
    prob = load_golden(name)
    fg, p, dim = int(prob["face_geom"]), int(prob["p"]), int(prob["dim"])
    q = p + 1
    a, b = (float(x) for x in prob["params"])
    face_table = fo.dof_info(fg, p, prob["face_native_dofs"])
    mesh_face_table = fo.dof_info(fg, int(prob["mesh_order"]), prob["mesh_face_native_dofs"])
    nodes_E = fo.gather(mesh_face_table, prob["mesh_nodes"], dim, int(prob["mesh_ordering"]) == 1)
    Xv, Xg = fo.interpolate(nodes_E, int(prob["mesh_order"]), q, dim - 1, fg)
    val, grad = fo.interpolate(fo.gather(face_table, prob["U"], dim), p, q, dim - 1, fg)
    if dim == 3:
        n = np.cross(Xg[:, :, 0, :], Xg[:, :, 1, :], axis=1)  # [nf][3][nq]
    else:  # segment faces: the length of the tangent
        n = Xg[:, :, 0, :]
    measure = np.linalg.norm(n, axis=1)
    f = a * (1.0 + Xv[:, 0, :]) * np.einsum("ecq,ecq->eq", val, val) + 0.5 * b * np.einsum("ecjq,ecjq->eq", grad, grad)
    q_E = np.einsum("eq,eq,q->e", f, measure, fo._face_weights(fg, q))
    assert rel_l2(q_E, prob["q_E"]) < 1e-13 and abs(q_E.sum() - prob["value"]) < 1e-13 * abs(prob["value"])


# ----------------------------------------------------------------------------------------------------------------------
# discontinuous (L2) parameter fields
# ----------------------------------------------------------------------------------------------------------------------
def _l2_names():
    import glob
    import os
    from harness import GOLDEN
    return sorted(os.path.splitext(os.path.basename(f))[0] for f in glob.glob(os.path.join(GOLDEN, "l2param_*.npz")))


@pytest.mark.parametrize("name", _l2_names())
def test_l2_parameter_field(name):
    """H1<p>(H1<p>, L2<0>): L2 dofs are lexicographic already (element_restriction.cpp:262-268), order 0 has one node per
    element; residual and Jacobian actions w.r.t. both arguments"""
    prob = load_golden(name)
    geom, p, dim = int(prob["geom"]), int(prob["p"]), int(prob["dim"])
    q = p + 1
    table = fo.dof_info(geom, p, prob["native_dofs"])
    l2_table = np.asarray(prob["l2_native_dofs"]).astype(np.uint64)  # no permutation
    assert np.array_equal(l2_table, prob["l2_dof_info"] & MASK)
    mesh_table = prob["mesh_dof_info"] & MASK
    X_q, J_q = fo.geometric_factors(fo.gather(mesh_table, prob["mesh_nodes"], dim, int(prob["mesh_ordering"]) == 1),
                                    int(prob["mesh_order"]), q, dim, geom)
    trials = [fo.gather(table, prob["U"], 1), fo.gather(l2_table, prob["K"], 1)]
    qf = fo.make_qf("param_diffusion", list(prob["params"]))
    r_E, _ = fo.evaluation_kernel(qf, trials, [p, 0], X_q, J_q, p, q, dim, geom=geom)
    assert rel_l2(r_E.ravel(), prob["r_E"]) < 1e-13
    for wrt, (tab, order, d) in enumerate(((table, p, prob["dU"]), (l2_table, 0, prob["dK"]))):
        _, blocks = fo.evaluation_kernel(qf, trials, [p, 0], X_q, J_q, p, q, dim, wrt=wrt, geom=geom)
        dr_E = fo.action_of_gradient_kernel(blocks, fo.gather(tab, d, 1), order, p, q, dim, geom)
        dr = fo.scatter_add(table, dr_E, np.zeros(len(prob["U"])), 1)
        assert rel_l2(dr, prob[f"dr{wrt}"]) < 1e-13
