"""TEST INFRASTRUCTURE ONLY (oracle/): builds parity problems with the REFERENCE oracle (oracle/ref.py) and writes them as
small golden fixtures under tests/golden/.  Each fixture holds the inputs the product needs (native element dofs, mesh
nodes, L-vectors) and the reference's outputs for them (dof_info, X_q/J_q, residual, JVP, element matrices, CSR).

    python oracle/make_golden.py            # regenerates every fixture listed in FIXTURES (needs oracle/_ref)

`reference_problem(...)` is also imported by the tests to run live comparisons at other sizes when oracle/_ref exists."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
REFERENCE_MESHES = "/root/reference/data/meshes"


def make_mesh(spec):
    """spec: ("file", name, refinements) | ("cart3d", nx, ny, nz, perturb) | ("cart2d", nx, ny, perturb)
    | ("cart3d_tet", nx, ny, nz, perturb) | ("cart2d_tri", nx, ny, perturb)"""
    if spec[0] == "file":
        return ref.Mesh.from_file(os.path.join(REFERENCE_MESHES, spec[1]), spec[2])
    if spec[0] == "cart3d":
        m = ref.Mesh.cartesian3d(spec[1], spec[2], spec[3])
        if spec[4]:
            m.perturb(spec[4])
        return m
    if spec[0] == "cart2d":
        m = ref.Mesh.cartesian2d(spec[1], spec[2])
        if spec[3]:
            m.perturb(spec[3])
        return m
    if spec[0] == "cart3d_tet":
        m = ref.Mesh.cartesian3d_tet(spec[1], spec[2], spec[3])
        if spec[4]:
            m.perturb(spec[4])
        return m
    if spec[0] == "cart2d_tri":
        m = ref.Mesh.cartesian2d_tri(spec[1], spec[2])
        if spec[3]:
            m.perturb(spec[3])
        return m
    raise ValueError(spec)


def reference_problem(mesh_spec, kind, p, params, vdim, seed=42, with_matrix=True, shape_amp=0.0, state_init=None,
                      u_amp=0.01, subset=None):
    """Runs the reference path for one Integral on the whole mesh with signature test(shape p=1 vector, u, a)
    (SolidMechanics: vdim = dim; HeatTransfer: vdim = 1) and returns inputs + outputs as a dict of numpy arrays."""
    mesh = make_mesh(mesh_spec)
    dim = mesh.dim
    tag = f"{dim}d_p{p}"
    space = ref.Space(mesh, p, vdim)
    shape_space = ref.Space(mesh, 1, dim)
    rng = np.random.default_rng(seed)
    U = u_amp * rng.uniform(-1.0, 1.0, space.vsize)  # small enough that F = I + du/dX stays invertible on this mesh
    A = rng.uniform(-1.0, 1.0, space.vsize)
    S = shape_amp * rng.uniform(-1.0, 1.0, shape_space.vsize)
    dU = rng.uniform(0.0, 1.0, space.vsize)  # check_gradient.hpp:19,26
    q = p + 1
    order, ordering, nodes, nodes_info = mesh.nodes()
    X_q, J_q = mesh.geometric_factors(q)

    # subset: callable ne -> bool mask; the integral then lives on Domain::ofElements of the selected elements
    # (F/tests/functional_with_domain.cpp): E-vectors stay full size, derivatives are indexed by position in the domain
    mask = None if subset is None else np.asarray(subset(mesh.ne), dtype=bool)
    I = ref.Integral(f"{kind}_{tag}", mesh, params, element_mask=mask)
    out = dict(
        kind=kind, geom=mesh.geom, dim=dim, p=p, vdim=vdim, ne=mesh.ne, params=np.asarray(params, dtype=np.float64),
        native_dofs=space.native_dofs(), dof_info=space.dof_info(), num_nodes=space.num_nodes,
        shape_native_dofs=shape_space.native_dofs(), shape_dof_info=shape_space.dof_info(),
        shape_num_nodes=shape_space.num_nodes,
        mesh_order=order, mesh_ordering=ordering, mesh_nodes=nodes, mesh_dof_info=nodes_info,
        mesh_num_nodes=len(nodes) // dim,
        U=U, A=A, S=S, dU=dU, X_q=X_q, J_q=J_q,
    )
    if mask is not None:
        out["domain_elements"] = np.nonzero(mask)[0].astype(np.int32)
    ins = [shape_space.gather(S), space.gather(U), space.gather(A)]
    out["U_E"] = ins[1]
    state_bytes = I.state_bytes()
    if state_bytes:
        if state_init is not None:
            I.set_qdata(state_init)
        out["qdata0"] = I.get_qdata()
    # residual without differentiation (functional.hpp:446-498 with P = identity)
    r_E = I.mult(0.0, ins, space.esize)
    out["r_E"] = r_E
    out["r"] = space.scatter_add(r_E, np.zeros(space.vsize))
    # residual with derivative store w.r.t. u (argument 1) -- values must equal the plain residual
    r_E_ad = I.mult(0.0, ins, space.esize, wrt=1)
    out["r_E_ad"] = r_E_ad
    # Jacobian action (functional.hpp:406-432)
    dr_E = I.jvp(1, space.gather(dU), space.esize)
    out["dr_E"] = dr_E
    out["dr"] = space.scatter_add(dr_E, np.zeros(space.vsize))
    # derivative w.r.t. the acceleration (argument 2: the mass matrix) and the shape field (argument 0)
    I.mult(0.0, ins, space.esize, wrt=2)
    out["dr_da"] = space.scatter_add(I.jvp(2, space.gather(dU), space.esize), np.zeros(space.vsize))
    dS = rng.uniform(0.0, 1.0, shape_space.vsize)
    out["dS"] = dS
    I.mult(0.0, ins, space.esize, wrt=0)
    out["dr_dshape"] = space.scatter_add(I.jvp(0, shape_space.gather(dS), space.esize), np.zeros(space.vsize))
    if state_bytes:
        # committing the state (domain_integral_kernels.hpp:143-147)
        I.mult(0.0, ins, space.esize, update_state=True)
        out["qdata1"] = I.get_qdata()
        out["r_E_after_commit"] = I.mult(0.0, ins, space.esize)
        I.set_qdata(out["qdata0"])
    if with_matrix:
        I.mult(0.0, ins, space.esize, wrt=1)
        npe = space.dof_info().shape[1]
        K_e = I.element_gradient(1, mesh.ne, npe * vdim, npe * vdim)
        out["K_e"] = K_e
        csr = ref.Csr(space, space)
        rp, ci = csr.pattern()
        out["row_ptr"], out["col_ind"] = rp, ci
        out["values"] = csr.assemble(K_e)
    return out


def reference_multifield_problem(mesh_spec, kind, p, params, test_vdim, trial_vdims, amps, seed=42, u_amp=0.01):
    """One reference domain Integral with an arbitrary list of trial fields after the shape displacement: signature
    H1<p,test_vdim>(shape H1<1,dim>, H1<p,trial_vdims[0]>, ...) -- the parameterized physics modules (thermomechanics:
    SolidMechanics<..., Parameters<H1<p>>> and HeatTransfer<..., Parameters<H1<p,dim>>>, thermomechanics.hpp:534-537).
    amps[i] scales the uniform(-1,1) draw of trial field i + 1 (offsets: see `offsets`).  Outputs: residual, the Jacobian
    action and the assembled (possibly rectangular) matrix w.r.t. EVERY argument, and the committed state."""
    mesh = make_mesh(mesh_spec)
    dim = mesh.dim
    tag = f"{dim}d_p{p}"
    test = ref.Space(mesh, p, test_vdim)
    shape_space = ref.Space(mesh, 1, dim)
    spaces, by_vdim = [shape_space], {}
    for v in trial_vdims:
        if v not in by_vdim:
            by_vdim[v] = test if v == test_vdim else ref.Space(mesh, p, v)
        spaces.append(by_vdim[v])
    rng = np.random.default_rng(seed)
    fields = [0.01 * rng.uniform(-1.0, 1.0, shape_space.vsize)]
    for sp, (amp, offset) in zip(spaces[1:], amps):
        fields.append(offset + amp * rng.uniform(-1.0, 1.0, sp.vsize))
    dfields = [rng.uniform(0.0, 1.0, sp.vsize) for sp in spaces]
    q = p + 1
    order, ordering, nodes, nodes_info = mesh.nodes()
    X_q, J_q = mesh.geometric_factors(q)
    I = ref.Integral(f"{kind}_{tag}", mesh, params)
    out = dict(kind=kind, geom=mesh.geom, dim=dim, p=p, ne=mesh.ne, params=np.asarray(params, dtype=np.float64),
               test_vdim=test_vdim, trial_vdims=np.asarray([dim] + list(trial_vdims)), num_nodes=test.num_nodes,
               native_dofs=test.native_dofs(), dof_info=test.dof_info(),
               shape_native_dofs=shape_space.native_dofs(), shape_num_nodes=shape_space.num_nodes,
               mesh_order=order, mesh_ordering=ordering, mesh_nodes=nodes, mesh_dof_info=nodes_info,
               mesh_num_nodes=len(nodes) // dim, X_q=X_q, J_q=J_q)
    for i, (f, df) in enumerate(zip(fields, dfields)):
        out[f"field{i}"], out[f"dfield{i}"] = f, df
    ins = [sp.gather(f) for sp, f in zip(spaces, fields)]
    if I.state_bytes():
        out["qdata0"] = I.get_qdata()
    r_E = I.mult(0.0, ins, test.esize)
    out["r_E"] = r_E
    out["r"] = test.scatter_add(r_E, np.zeros(test.vsize))
    npe = test.dof_info().shape[1]
    for i, sp in enumerate(spaces):
        r_E_ad = I.mult(0.0, ins, test.esize, wrt=i)
        assert np.array_equal(r_E_ad, r_E) or np.allclose(r_E_ad, r_E, rtol=1e-13, atol=1e-300)
        out[f"dr{i}"] = test.scatter_add(I.jvp(i, sp.gather(dfields[i]), test.esize), np.zeros(test.vsize))
        npe_k = sp.dof_info().shape[1]
        vk = dim if i == 0 else trial_vdims[i - 1]
        K_e = I.element_gradient(i, mesh.ne, npe_k * vk, npe * test_vdim)
        csr = ref.Csr(test, sp)
        rp, ci = csr.pattern()
        out[f"row_ptr{i}"], out[f"col_ind{i}"], out[f"values{i}"] = rp, ci, csr.assemble(K_e)
    if I.state_bytes():
        I.mult(0.0, ins, test.esize, update_state=True)
        out["qdata1"] = I.get_qdata()
        out["r_E_after_commit"] = I.mult(0.0, ins, test.esize)
    return out


# thermomechanics (BASELINE.json configs[3]: coupled HeatTransfer + SolidMechanics, thermoelastic material, tet mesh, p=1)
# params: density E nu C_v alpha theta_ref k [geom_nonlin]; amps = (amplitude, offset) per trial field after the shape
_GSV = [1.0, 100.0, 0.25, 1.0, 1.0e-3, 1.0, 1.0]
MULTIFIELD_FIXTURES = [
    ("thermoelastic_solid_tet_p1_beam", dict(mesh_spec=("file", "beam-tet.mesh", 0), kind="thermoelastic_solid", p=1,
                                             params=_GSV + [1.0], test_vdim=3, trial_vdims=[3, 3, 1],
                                             amps=[(0.01, 0.0), (1.0, 0.0), (0.5, 2.0)])),
    ("thermoelastic_thermal_tet_p1_beam", dict(mesh_spec=("file", "beam-tet.mesh", 0), kind="thermoelastic_thermal",
                                               p=1, params=_GSV, test_vdim=1, trial_vdims=[1, 1, 3],
                                               amps=[(0.5, 2.0), (1.0, 0.0), (0.01, 0.0)])),
    ("thermoelastic_solid_hex_p2_cart", dict(mesh_spec=("cart3d", 2, 1, 1, 0.02), kind="thermoelastic_solid", p=2,
                                             params=_GSV + [1.0], test_vdim=3, trial_vdims=[3, 3, 1],
                                             amps=[(0.01, 0.0), (1.0, 0.0), (0.5, 2.0)])),
    ("thermoelastic_thermal_tet_p2_cart", dict(mesh_spec=("cart3d_tet", 1, 1, 1, 0.02), kind="thermoelastic_thermal",
                                               p=2, params=_GSV, test_vdim=1, trial_vdims=[1, 1, 3],
                                               amps=[(0.5, 2.0), (1.0, 0.0), (0.01, 0.0)])),
]


def reference_qoi_problem(mesh_spec, kind, p, params, seed=42):
    """One reference QOI Integral (Functional<double(H1<p,dim>)>, functional_qoi.inl): the per-element totals, the
    scalar, its directional derivative and its gradient vector (assemble(grad) of a QOI is a vector:
    functional_qoi.inl:411-455)."""
    mesh = make_mesh(mesh_spec)
    dim = mesh.dim
    space = ref.Space(mesh, p, dim)
    rng = np.random.default_rng(seed)
    U = rng.uniform(-1.0, 1.0, space.vsize)
    dU = rng.uniform(0.0, 1.0, space.vsize)
    order, ordering, nodes, nodes_info = mesh.nodes()
    X_q, J_q = mesh.geometric_factors(p + 1)
    I = ref.Integral(f"{kind}_{dim}d_p{p}", mesh, params)
    out = dict(kind=kind, geom=mesh.geom, dim=dim, p=p, ne=mesh.ne, params=np.asarray(params, dtype=np.float64),
               native_dofs=space.native_dofs(), dof_info=space.dof_info(), num_nodes=space.num_nodes,
               mesh_order=order, mesh_ordering=ordering, mesh_nodes=nodes, mesh_dof_info=nodes_info,
               mesh_num_nodes=len(nodes) // dim, U=U, dU=dU, X_q=X_q, J_q=J_q)
    ins = [space.gather(U)]
    q_E = I.mult(0.0, ins, mesh.ne)  # one total per element (QoIElementRestriction: ESize = number of elements)
    out["q_E"] = q_E
    out["value"] = float(np.sum(q_E))
    I.mult(0.0, ins, mesh.ne, wrt=0)
    out["dq_E"] = I.jvp(0, space.gather(dU), mesh.ne)
    out["dvalue"] = float(np.sum(out["dq_E"]))
    npe = space.dof_info().shape[1]
    K_e = I.element_gradient(0, mesh.ne, npe * dim, 1)  # [element][trial vdof][1]
    out["K_e"] = K_e
    out["gradient"] = space.scatter_add(K_e.reshape(mesh.ne, dim, npe), np.zeros(space.vsize))
    return out


def reference_l2_parameter_problem(mesh_spec, p, params, seed=42):
    """H1<p>(H1<p> u, L2<0> k): a domain integral with a discontinuous piecewise-constant parameter field (the L2 dofs mfem
    returns are lexicographic already: no permutation, element_restriction.cpp:262-268): residual, the Jacobian action
    and the assembled matrix w.r.t. both arguments (the second one is rectangular: H1 rows x L2 columns)."""
    mesh = make_mesh(mesh_spec)
    dim = mesh.dim
    space = ref.Space(mesh, p, 1)
    pspace = ref.Space(mesh, 0, 1, family="L2")
    rng = np.random.default_rng(seed)
    U, K = rng.uniform(-1.0, 1.0, space.vsize), rng.uniform(0.5, 1.5, pspace.vsize)
    dU, dK = rng.uniform(0.0, 1.0, space.vsize), rng.uniform(0.0, 1.0, pspace.vsize)
    order, ordering, nodes, nodes_info = mesh.nodes()
    X_q, J_q = mesh.geometric_factors(p + 1)
    I = ref.Integral(f"param_diffusion_{dim}d_p{p}", mesh, params)
    out = dict(kind="param_diffusion", geom=mesh.geom, dim=dim, p=p, ne=mesh.ne, params=np.asarray(params, dtype=np.float64),
               native_dofs=space.native_dofs(), dof_info=space.dof_info(), num_nodes=space.num_nodes,
               l2_native_dofs=pspace.native_dofs(), l2_dof_info=pspace.dof_info(), l2_num_nodes=pspace.num_nodes,
               mesh_order=order, mesh_ordering=ordering, mesh_nodes=nodes, mesh_dof_info=nodes_info,
               mesh_num_nodes=len(nodes) // dim, U=U, K=K, dU=dU, dK=dK, X_q=X_q, J_q=J_q)
    ins = [space.gather(U), pspace.gather(K)]
    r_E = I.mult(0.0, ins, space.esize)
    out["r_E"] = r_E
    out["r"] = space.scatter_add(r_E, np.zeros(space.vsize))
    npe = space.dof_info().shape[1]
    for i, (sp, d) in enumerate(((space, dU), (pspace, dK))):
        I.mult(0.0, ins, space.esize, wrt=i)
        out[f"dr{i}"] = space.scatter_add(I.jvp(i, sp.gather(d), space.esize), np.zeros(space.vsize))
        K_e = I.element_gradient(i, mesh.ne, sp.dof_info().shape[1], npe)
        csr = ref.Csr(space, sp)
        rp, ci = csr.pattern()
        out[f"row_ptr{i}"], out[f"col_ind{i}"], out[f"values{i}"] = rp, ci, csr.assemble(K_e)
    return out


L2_FIXTURES = [
    ("l2param_diffusion_hex_p2_cart", dict(mesh_spec=("cart3d", 2, 2, 1, 0.02), p=2, params=[0.3])),
    ("l2param_diffusion_tet_p1_patch", dict(mesh_spec=("file", "patch3D_tets.mesh", 0), p=1, params=[0.3])),
]

def reference_boundary_qoi_problem(mesh_spec, kind, p, params, seed=42):
    """One reference BOUNDARY QOI Integral on EntireBoundary(mesh): per-face totals, the scalar, its directional
    derivative and its gradient vector."""
    mesh = make_mesh(mesh_spec)
    dim = mesh.dim
    space = ref.Space(mesh, p, dim)
    fspace = ref.FaceSpace(space)
    rng = np.random.default_rng(seed)
    U = rng.uniform(-1.0, 1.0, space.vsize)
    dU = rng.uniform(0.0, 1.0, space.vsize)
    order, ordering, nodes, _ = mesh.nodes()
    mesh_face_native, mesh_face_info = mesh.nodes_face_tables()
    X_q, J_q = mesh.boundary_geometric_factors(p + 1)
    I = ref.Integral(f"{kind}_{dim}d_p{p}", mesh, params, boundary=True)
    nf = fspace.nf
    out = dict(kind=kind, geom=mesh.geom, face_geom=mesh.face_geom, dim=dim, p=p, ne=mesh.ne, nf=nf,
               params=np.asarray(params, dtype=np.float64), native_dofs=space.native_dofs(), num_nodes=space.num_nodes,
               face_native_dofs=fspace.native_dofs(), face_dof_info=fspace.dof_info(),
               mesh_order=order, mesh_ordering=ordering, mesh_nodes=nodes, mesh_num_nodes=len(nodes) // dim,
               mesh_face_native_dofs=mesh_face_native, mesh_face_dof_info=mesh_face_info, U=U, dU=dU, X_q=X_q, J_q=J_q)
    ins = [fspace.gather(U)]
    q_E = I.mult(0.0, ins, nf)
    out["q_E"] = q_E
    out["value"] = float(np.sum(q_E))
    I.mult(0.0, ins, nf, wrt=0)
    out["dvalue"] = float(np.sum(I.jvp(0, fspace.gather(dU), nf)))
    npf = fspace.dof_info().shape[1]
    K_e = I.element_gradient(0, nf, npf * dim, 1)
    out["gradient"] = fspace.scatter_add(K_e.reshape(nf, dim, npf), np.zeros(space.vsize))
    return out


BOUNDARY_QOI_FIXTURES = [
    ("qoi_boundary_quadratic_hex_p2_cart", dict(mesh_spec=("cart3d", 2, 1, 2, 0.02), kind="qoi_boundary_quadratic", p=2,
                                                params=[0.7, 1.3])),
    ("qoi_boundary_quadratic_tet_p1_beam", dict(mesh_spec=("file", "beam-tet.mesh", 0), kind="qoi_boundary_quadratic",
                                                p=1, params=[0.7, 1.3])),
    # segment faces of 2-D meshes (F/detail/segment_H1.inl; F/tests/bug_boundary_qoi.cpp runs on a 2-D mesh)
    ("qoi_boundary_quadratic_quad_p2_cart", dict(mesh_spec=("cart2d", 3, 2, 0.02), kind="qoi_boundary_quadratic", p=2,
                                                 params=[0.7, 1.3])),
    ("qoi_boundary_quadratic_tri_p1_patch", dict(mesh_spec=("file", "patch2D_tris.mesh", 1),
                                                 kind="qoi_boundary_quadratic", p=1, params=[0.7, 1.3])),
]

QOI_FIXTURES = [
    ("qoi_quadratic_hex_p2_cart", dict(mesh_spec=("cart3d", 2, 2, 1, 0.02), kind="qoi_quadratic", p=2, params=[0.7, 1.3])),
    ("qoi_quadratic_tet_p1_beam", dict(mesh_spec=("file", "beam-tet.mesh", 0), kind="qoi_quadratic", p=1,
                                       params=[0.7, 1.3])),
]


def reference_boundary_problem(mesh_spec, kind, p, params, seed=42, shape_amp=0.0, u_amp=0.01):
    """One reference BOUNDARY Integral (MakeBoundaryIntegral, integral.hpp:322-350) on EntireBoundary(mesh) with the
    SolidMechanics signature H1<p,dim>(shape H1<1,dim>, u, a): inputs (native face dofs of every space, mesh nodes) and
    the reference's outputs (face dof_info, face geometric factors, residual, JVPs, face element matrices and their
    contribution to the CSR values of the domain pattern)."""
    mesh = make_mesh(mesh_spec)
    dim = mesh.dim
    tag = f"{dim}d_p{p}"
    space, shape_space = ref.Space(mesh, p, dim), ref.Space(mesh, 1, dim)
    fspace, fshape = ref.FaceSpace(space), ref.FaceSpace(shape_space)
    rng = np.random.default_rng(seed)
    U = u_amp * rng.uniform(-1.0, 1.0, space.vsize)
    A = rng.uniform(-1.0, 1.0, space.vsize)
    S = shape_amp * rng.uniform(-1.0, 1.0, shape_space.vsize)
    dU = rng.uniform(0.0, 1.0, space.vsize)
    dS = rng.uniform(0.0, 1.0, shape_space.vsize)
    q = p + 1
    order, ordering, nodes, nodes_info = mesh.nodes()
    mesh_face_native, mesh_face_info = mesh.nodes_face_tables()
    X_q, J_q = mesh.boundary_geometric_factors(q)
    I = ref.Integral(f"{kind}_{tag}", mesh, params, boundary=True)
    out = dict(
        kind=kind, geom=mesh.geom, face_geom=mesh.face_geom, dim=dim, p=p, vdim=dim, ne=mesh.ne, nf=fspace.nf,
        params=np.asarray(params, dtype=np.float64),
        native_dofs=space.native_dofs(), num_nodes=space.num_nodes, shape_num_nodes=shape_space.num_nodes,
        face_native_dofs=fspace.native_dofs(), face_dof_info=fspace.dof_info(),
        shape_face_native_dofs=fshape.native_dofs(), shape_face_dof_info=fshape.dof_info(),
        mesh_order=order, mesh_ordering=ordering, mesh_nodes=nodes, mesh_num_nodes=len(nodes) // dim,
        mesh_face_native_dofs=mesh_face_native, mesh_face_dof_info=mesh_face_info,
        U=U, A=A, S=S, dU=dU, dS=dS, X_q=X_q, J_q=J_q,
    )
    ins = [fshape.gather(S), fspace.gather(U), fspace.gather(A)]
    r_E = I.mult(0.0, ins, fspace.esize)
    out["r_E"] = r_E
    out["r"] = fspace.scatter_add(r_E, np.zeros(space.vsize))
    out["r_E_ad"] = I.mult(0.0, ins, fspace.esize, wrt=1)
    out["dr"] = fspace.scatter_add(I.jvp(1, fspace.gather(dU), fspace.esize), np.zeros(space.vsize))
    I.mult(0.0, ins, fspace.esize, wrt=0)
    out["dr_dshape"] = fspace.scatter_add(I.jvp(0, fshape.gather(dS), fspace.esize), np.zeros(space.vsize))
    I.mult(0.0, ins, fspace.esize, wrt=1)
    npe = fspace.dof_info().shape[1]
    K_e = I.element_gradient(1, fspace.nf, npe * dim, npe * dim)
    out["K_e"] = K_e
    csr = ref.Csr(space, space)
    rp, ci = csr.pattern()
    out["row_ptr"], out["col_ind"] = rp, ci
    out["values"] = csr.assemble_faces(K_e, np.zeros(csr.nnz))
    return out


BOUNDARY_FIXTURES = [
    # follower pressure on the current configuration: depends on u, so d/du is non-zero (setPressure)
    ("solid_pressure_hex_p2_cart", dict(mesh_spec=("cart3d", 2, 1, 2, 0.02), kind="solid_pressure", p=2,
                                        params=[3.0, 1.0], shape_amp=0.01)),
    ("solid_pressure_hex_p1_patch", dict(mesh_spec=("file", "patch3D_hexes.mesh", 0), kind="solid_pressure", p=1,
                                         params=[-1.5, 1.0], shape_amp=0.01)),
    # triangular faces of tet meshes (BASELINE.json configs[3] runs on a tet mesh)
    ("solid_pressure_tet_p2_cart", dict(mesh_spec=("cart3d_tet", 1, 1, 2, 0.02), kind="solid_pressure", p=2,
                                        params=[3.0, 1.0], shape_amp=0.01)),
    ("solid_traction_tet_p1_beam", dict(mesh_spec=("file", "beam-tet.mesh", 0), kind="solid_traction", p=1,
                                        params=[0.3, -1.0, 2.0], shape_amp=0.01)),
    # dead traction: only the shape derivative is non-zero (setTraction)
    ("solid_traction_hex_p2_cart", dict(mesh_spec=("cart3d", 2, 2, 1, 0.02), kind="solid_traction", p=2,
                                        params=[0.3, -1.0, 2.0], shape_amp=0.01)),
    # segment faces of 2-D meshes (F/detail/segment_H1.inl): quadrilateral and triangular meshes, p = 1 and 2;
    # star.mesh is the mesh of BASELINE.json configs[0]
    ("solid_pressure_quad_p2_cart", dict(mesh_spec=("cart2d", 3, 2, 0.02), kind="solid_pressure", p=2,
                                         params=[3.0, 1.0], shape_amp=0.01)),
    ("solid_traction_quad_p1_star", dict(mesh_spec=("file", "star.mesh", 0), kind="solid_traction", p=1,
                                         params=[0.3, -1.0], shape_amp=0.01)),
    ("solid_pressure_tri_p2_cart", dict(mesh_spec=("cart2d_tri", 3, 2, 0.02), kind="solid_pressure", p=2,
                                        params=[-1.5, 1.0], shape_amp=0.01)),
    ("solid_traction_tri_p1_patch", dict(mesh_spec=("file", "patch2D_tris.mesh", 1), kind="solid_traction", p=1,
                                         params=[0.3, -1.0], shape_amp=0.01)),
]

# (file name, kwargs of reference_problem)
FIXTURES = [
    ("solid_neohookean_hex_p2_patch", dict(mesh_spec=("file", "patch3D_hexes.mesh", 0), kind="solid_neohookean", p=2,
                                           params=[1.0, 100.0, 50.0, 1.0], vdim=3, shape_amp=0.01)),
    ("solid_neohookean_hex_p1_beam", dict(mesh_spec=("file", "beam-hex.mesh", 0), kind="solid_neohookean", p=1,
                                          params=[1.0, 100.0, 50.0, 1.0], vdim=3)),
    ("solid_neohookean_hex_p3_cart", dict(mesh_spec=("cart3d", 2, 2, 1, 0.02), kind="solid_neohookean", p=3,
                                          params=[1.0, 100.0, 50.0, 1.0], vdim=3, with_matrix=False)),
    ("solid_linear_isotropic_hex_p2_cart", dict(mesh_spec=("cart3d", 2, 1, 2, 0.02), kind="solid_linear_isotropic",
                                                p=2, params=[2.0, 100.0, 50.0, 0.0], vdim=3)),
    ("solid_neohookean_quad_p2_cart", dict(mesh_spec=("cart2d", 3, 2, 0.02), kind="solid_neohookean", p=2,
                                           params=[1.0, 100.0, 50.0, 1.0], vdim=2)),
    # BASELINE.json configs[2]: finite-deformation J2 (logarithmic strain, exponential return map) with the inverse
    # plastic distortion + eqps as QuadratureData state; E nu sigma_y Hi density geom_nonlin; plastic at most points
    ("solid_j2_linear_hex_p2_cart", dict(mesh_spec=("cart3d", 2, 1, 2, 0.02), kind="solid_j2_linear", p=2,
                                         params=[100.0, 0.25, 0.5, 5.0, 1.0, 1.0], vdim=3, u_amp=0.02)),
    # integrals on part of the mesh (Domain::ofElements; F/tests/functional_with_domain.cpp)
    ("solid_neohookean_hex_p2_subdomain", dict(mesh_spec=("cart3d", 3, 2, 2, 0.02), kind="solid_neohookean", p=2,
                                               params=[1.0, 100.0, 50.0, 1.0], vdim=3, shape_amp=0.01,
                                               subset=lambda ne: np.arange(ne) % 3 != 1)),
    ("thermal_linear_conductor_tet_p2_subdomain", dict(mesh_spec=("cart3d_tet", 2, 1, 1, 0.02),
                                                       kind="thermal_linear_conductor", p=2, params=[1.5, 2.0, 0.7],
                                                       vdim=1, subset=lambda ne: np.arange(ne) % 4 != 2)),
    # simplex elements (BASELINE.json configs[3] runs on a tet mesh, H1 p=1)
    ("solid_neohookean_tet_p1_beam", dict(mesh_spec=("file", "beam-tet.mesh", 0), kind="solid_neohookean", p=1,
                                          params=[1.0, 100.0, 50.0, 1.0], vdim=3, shape_amp=0.01)),
    ("solid_neohookean_tet_p2_cart", dict(mesh_spec=("cart3d_tet", 2, 1, 1, 0.02), kind="solid_neohookean", p=2,
                                          params=[1.0, 100.0, 50.0, 1.0], vdim=3, shape_amp=0.01)),
    ("solid_linear_isotropic_tet_p3_patch", dict(mesh_spec=("file", "patch3D_tets.mesh", 0),
                                                 kind="solid_linear_isotropic", p=3, params=[2.0, 100.0, 50.0, 0.0],
                                                 vdim=3, with_matrix=False)),
    ("solid_neohookean_tri_p2_cart", dict(mesh_spec=("cart2d_tri", 3, 2, 0.02), kind="solid_neohookean", p=2,
                                          params=[1.0, 100.0, 50.0, 1.0], vdim=2, shape_amp=0.01)),
    ("thermal_linear_conductor_tet_p1_patch", dict(mesh_spec=("file", "patch3D_tets.mesh", 0),
                                                   kind="thermal_linear_conductor", p=1, params=[1.5, 2.0, 0.7],
                                                   vdim=1, shape_amp=0.01)),
    ("thermal_linear_conductor_tri_p1_patch", dict(mesh_spec=("file", "patch2D_tris.mesh", 1),
                                                   kind="thermal_linear_conductor", p=1, params=[1.5, 2.0, 0.7],
                                                   vdim=1)),
    # BASELINE.json configs[0]: HeatTransfer linear conduction, H1 p=1 on star.mesh (20 quads)
    ("thermal_linear_conductor_quad_p1_star", dict(mesh_spec=("file", "star.mesh", 1), kind="thermal_linear_conductor",
                                                   p=1, params=[1.5, 2.0, 0.7], vdim=1, shape_amp=0.01)),
    ("thermal_linear_conductor_hex_p2_cart", dict(mesh_spec=("cart3d", 2, 2, 1, 0.02), kind="thermal_linear_conductor",
                                                  p=2, params=[1.0, 1.0, 3.0], vdim=1)),
    # BASELINE.json configs[2] material family: J2 plasticity with QuadratureData state (plastic loading: the return map
    # runs at most points with these parameters), E nu sigma_y Hi Hk density geom_nonlin
    # the remaining hardening laws of solid_material.hpp (:154-198) and St. Venant-Kirchhoff (:61-92)
    ("solid_j2smallstrain_powerlaw_hex_p2_cart", dict(mesh_spec=("cart3d", 2, 1, 2, 0.02),
                                                      kind="solid_j2smallstrain_powerlaw", p=2,
                                                      params=[100.0, 0.25, 0.5, 3.0, 0.01, 2.0, 1.0, 0.0], vdim=3,
                                                      u_amp=0.02)),
    ("solid_j2smallstrain_voce_hex_p2_cart", dict(mesh_spec=("cart3d", 2, 1, 2, 0.02),
                                                  kind="solid_j2smallstrain_voce", p=2,
                                                  params=[100.0, 0.25, 0.5, 0.9, 0.02, 2.0, 1.0, 0.0], vdim=3,
                                                  u_amp=0.02)),
    ("solid_stvk_hex_p2_cart", dict(mesh_spec=("cart3d", 2, 1, 2, 0.02), kind="solid_stvk", p=2,
                                    params=[1.0, 100.0, 50.0, 1.0], vdim=3)),
    ("solid_j2smallstrain_linear_hex_p2_cart", dict(mesh_spec=("cart3d", 2, 1, 2, 0.02),
                                                    kind="solid_j2smallstrain_linear", p=2,
                                                    params=[100.0, 0.25, 0.5, 5.0, 2.0, 1.0, 0.0], vdim=3, u_amp=0.02)),
]


def reference_l2_face_tables():
    """DG face restrictions (element_restriction.cpp:66-215, 304-355): for several meshes and orders the per-(face,
    element) inputs mfem provides and the reference's FaceType::BOUNDARY / FaceType::INTERIOR dof tables."""
    out, cases = {}, []
    specs = [("hex", ("cart3d", 2, 2, 2, 0.02)), ("tet", ("cart3d_tet", 2, 2, 1, 0.02)), ("quad", ("cart2d", 3, 2, 0.02)),
             ("tri", ("cart2d_tri", 3, 2, 0.02)), ("beamtet", ("file", "beam-tet.mesh", 0)),
             ("star", ("file", "star.mesh", 0))]
    for tag, spec in specs:
        mesh = make_mesh(spec)
        for p in (0, 1, 2, 3):
            space = ref.Space(mesh, p, 1, family="L2")
            for interior in (False, True):
                info, ins = ref.l2_face_tables(space, mesh.face_geom, interior)
                if len(ins["side"]) == 0:
                    continue
                name = f"{tag}_p{p}_{'interior' if interior else 'boundary'}"
                cases.append(name)
                out[name + "_face_geom"] = mesh.face_geom
                out[name + "_p"] = p
                out[name + "_dof_info"] = info
                for k, v in ins.items():
                    out[f"{name}_{k}"] = v
    out["cases"] = np.array(cases)
    return out


def config_a_fixture():
    """BASELINE.json configs[1]: beam-hex.mesh refined 4x (32 768 hexes), H1 p=2 vdim 3 (839 619 vdofs).  Stores the
    host-library inputs (native dofs, mesh nodes) and a strided sample + norms of the reference residual / Jacobian
    action so the full-size GPU run can be checked without shipping 13 MB of doubles."""
    kw = dict(mesh_spec=("file", "beam-hex.mesh", 4), kind="solid_neohookean", p=2, params=[1.0, 100.0, 50.0, 1.0],
              vdim=3, with_matrix=False, u_amp=1e-3)  # h = 1/16: 0.01 would fold elements (det F < 0 -> NaN)
    d = reference_problem(**kw)
    assert np.isfinite(d["r"]).all() and np.isfinite(d["dr"]).all()
    stride = 97
    keep = dict(kind=d["kind"], geom=d["geom"], dim=d["dim"], p=d["p"], vdim=d["vdim"], ne=d["ne"], params=d["params"],
                native_dofs=d["native_dofs"].astype(np.int32), num_nodes=d["num_nodes"],
                shape_native_dofs=d["shape_native_dofs"].astype(np.int32), shape_num_nodes=d["shape_num_nodes"],
                mesh_order=d["mesh_order"], mesh_ordering=d["mesh_ordering"], mesh_nodes=d["mesh_nodes"],
                mesh_num_nodes=d["mesh_num_nodes"], seed=42, u_amp=kw["u_amp"], sample_stride=stride,
                r_sample=d["r"][::stride], dr_sample=d["dr"][::stride], r_norm=np.linalg.norm(d["r"]),
                dr_norm=np.linalg.norm(d["dr"]), r_sum=d["r"].sum(), dr_sum=d["dr"].sum())
    return keep


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    only = sys.argv[1:]
    if not only or "config_a" in only:
        data = config_a_fixture()
        path = os.path.join(GOLDEN_DIR, "config_a_beam_hex_r4_p2.npz")
        np.savez_compressed(path, **data)
        print(f"config A: ne={data['ne']} -> {os.path.getsize(path) / 1024:.0f} KiB")
    if not only or "l2faces" in only:
        data = reference_l2_face_tables()
        path = os.path.join(GOLDEN_DIR, "l2faces_tables.npz")
        np.savez_compressed(path, **data)
        print(f"l2faces: {len(data['cases'])} cases -> {os.path.getsize(path) / 1024:.0f} KiB")
    for name, kw in FIXTURES:
        if only and name not in only:
            continue
        data = reference_problem(**kw)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **data)
        print(f"{name}: ne={data['ne']} vsize={len(data['U'])} -> {os.path.getsize(path) / 1024:.0f} KiB")
    for name, kw in MULTIFIELD_FIXTURES:
        if only and name not in only:
            continue
        data = reference_multifield_problem(**kw)
        path = os.path.join(GOLDEN_DIR, "multi_" + name + ".npz")
        np.savez_compressed(path, **data)
        print(f"{name}: ne={data['ne']} |r|={np.linalg.norm(data['r']):.3e} " +
              " ".join(f"|dr{i}|={np.linalg.norm(data[f'dr{i}']):.3e}" for i in range(4)) +
              f" -> {os.path.getsize(path) / 1024:.0f} KiB")
    for name, kw in QOI_FIXTURES:
        if only and name not in only:
            continue
        data = reference_qoi_problem(**kw)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **data)
        print(f"{name}: ne={data['ne']} value={data['value']:.12e} dvalue={data['dvalue']:.6e} "
              f"|gradient|={np.linalg.norm(data['gradient']):.6e} -> {os.path.getsize(path) / 1024:.0f} KiB")
    for name, kw in L2_FIXTURES:
        if only and name not in only:
            continue
        data = reference_l2_parameter_problem(**kw)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **data)
        print(f"{name}: ne={data['ne']} |r|={np.linalg.norm(data['r']):.6e} |dr0|={np.linalg.norm(data['dr0']):.6e} "
              f"|dr1|={np.linalg.norm(data['dr1']):.6e} -> {os.path.getsize(path) / 1024:.0f} KiB")
    for name, kw in BOUNDARY_QOI_FIXTURES:
        if only and name not in only:
            continue
        data = reference_boundary_qoi_problem(**kw)
        path = os.path.join(GOLDEN_DIR, "bqoi_" + name[len("qoi_boundary_"):] + ".npz")
        np.savez_compressed(path, **data)
        print(f"{name}: faces={data['nf']} value={data['value']:.12e} dvalue={data['dvalue']:.6e} "
              f"|gradient|={np.linalg.norm(data['gradient']):.6e} -> {os.path.getsize(path) / 1024:.0f} KiB")
    for name, kw in BOUNDARY_FIXTURES:
        if only and name not in only:
            continue
        data = reference_boundary_problem(**kw)
        path = os.path.join(GOLDEN_DIR, "bdr_" + name + ".npz")
        np.savez_compressed(path, **data)
        print(f"{name}: faces={data['nf']} vsize={len(data['U'])} |r|={np.linalg.norm(data['r']):.3e} "
              f"|dr|={np.linalg.norm(data['dr']):.3e} |dr_dshape|={np.linalg.norm(data['dr_dshape']):.3e} -> "
              f"{os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
