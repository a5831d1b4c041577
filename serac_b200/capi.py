"""ctypes binding of libserac_b200.so (include/serac_b200.h) plus thin Python mirrors of the reference objects the C ABI
replaces (ElementRestriction, Integral, Functional).  PyTorch is used only for device memory and streams.

The library is the product: if it is missing or a call fails this module raises -- there is no CPU fallback."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libserac_b200.so")

SEGMENT, TRIANGLE, SQUARE, TETRAHEDRON, CUBE = 1, 2, 3, 4, 5
BYNODES, BYVDIM = 0, 1
NO_DIFFERENTIATION = -1
MAX_TRIALS = 6

# every symbol include/serac_b200.h declares: (restype, argtypes)
_vp, _i, _d, _i64, _sz = C.c_void_p, C.c_int, C.c_double, C.c_int64, C.c_size_t
_pp = C.POINTER(C.c_void_p)


class FieldArg(C.Structure):
    _fields_ = [("L_dev", _vp), ("table_dev", _vp), ("node_stride", _i64), ("comp_stride", _i64)]


class KernelTable(C.Structure):
    _fields_ = [
        ("geom", _i), ("dim", _i), ("q", _i), ("num_trials", _i), ("test_order", _i), ("test_components", _i),
        ("trial_order", _i * MAX_TRIALS), ("trial_components", _i * MAX_TRIALS),
        ("qf_bytes", _sz), ("state_bytes", _sz), ("elements_per_block", _i),
        ("evaluate", _vp), ("evaluate_ad", _vp * MAX_TRIALS), ("jvp", _vp * MAX_TRIALS),
        ("element_gradient", _vp * MAX_TRIALS),
        ("derivative_mask", _i * MAX_TRIALS), ("derivative_entries", _i * MAX_TRIALS), ("spatial_dim", _i),
        ("assemble_rows", _vp * MAX_TRIALS), ("assemble_rows_fixed_doubles", _i * MAX_TRIALS),
        ("assemble_rows_trial_blocks", _i * MAX_TRIALS), ("derivative_tile_elements", _i),
        ("derivative_full_entries", _i * MAX_TRIALS), ("specialize", _vp),
    ]


SIGNATURES = {
    "sb_version": (_i, []),
    "sb_last_error": (C.c_char_p, []),
    "sb_launch_count": (C.c_longlong, []),
    "sb_launch_count_reset": (None, []),
    "sb_count_launch": (None, []),
    "sb_h1_lexicographic_dof_info": (_i, [_i, _i, _vp, _i64, _vp]),
    "sb_l2_face_dof_info": (_i, [_i, _i, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sb_h1_lexicographic_permutation": (_i, [_i, _i, _vp]),
    "sb_restriction_create": (_i, [_vp, _i64, _i, _i, _i64, _i, _pp]),
    "sb_restriction_destroy": (_i, [_vp]),
    "sb_restriction_esize": (_i64, [_vp]),
    "sb_restriction_lsize": (_i64, [_vp]),
    "sb_restriction_num_elements": (_i64, [_vp]),
    "sb_restriction_table_dev": (_i, [_vp, _pp]),
    "sb_restriction_gather": (_i, [_vp, _vp, _vp, _vp]),
    "sb_restriction_scatter_add": (_i, [_vp, _vp, _vp, _vp]),
    "sb_restriction_scatter": (_i, [_vp, _vp, _vp, _vp]),
    "sb_geometric_factors": (_i, [_i, _i, _i, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "sb_named_integrand_count": (_i, []),
    "sb_named_integrand_name": (C.c_char_p, [_i]),
    "sb_named_kernel_table": (_i, [C.c_char_p, C.POINTER(C.POINTER(KernelTable))]),
    "sb_named_qf_blob": (_i, [C.c_char_p, _vp, _i, _vp, _sz]),
    "sb_integral_create": (_i, [C.POINTER(KernelTable), _vp, _vp, _i64, _vp, _vp, _pp]),
    "sb_integral_destroy": (_i, [_vp]),
    "sb_integral_set_qdata": (_i, [_vp, _vp]),
    "sb_integral_evaluate": (_i, [_vp, _d, C.POINTER(FieldArg), _vp, _i, _i, _vp]),
    "sb_integral_jvp": (_i, [_vp, _i, C.POINTER(FieldArg), _vp, _vp]),
    "sb_integral_evaluate_layout": (_i, [_vp, _d, C.POINTER(FieldArg), _vp, _i, _i, _i, _vp]),
    "sb_integral_jvp_layout": (_i, [_vp, _i, C.POINTER(FieldArg), _vp, _i, _vp]),
    "sb_restriction_scatter_layout": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "sb_integral_element_gradient": (_i, [_vp, _i, _vp, _vp]),
    "sb_integral_export_derivatives": (_i, [_vp, _i, _vp, _vp]),
    "sb_integral_derivative_entries": (_i, [_vp, _i]),
    "sb_integral_derivative_stored_entries": (_i, [_vp, _i]),
    "sb_integral_derivative_mask": (_i, [_vp, _i]),
    "sb_csr_create": (_i, [_vp, _vp, _pp]),
    "sb_csr_create_blocks": (_i, [_i, _vp, _vp, _pp]),
    "sb_csr_transpose": (_i, [_vp, _pp]),
    "sb_csr_transpose_values": (_i, [_vp, _vp, _vp, _vp]),
    "sb_csr_pattern_host": (_i, [_vp, _i, _i, _i64, _i, _vp, _i, _i, _i64, _i, _i64, _vp, _vp, C.POINTER(_i64)]),
    "sb_debug_assembly_plan_host": (_i, [_vp, _i, _i, _i64, _i, _vp, _i, _i, _i64, _i, _i64, _vp, _i64, _i, _i, _vp, _vp,
                                        _vp, C.POINTER(_i)]),
    "sb_csr_destroy": (_i, [_vp]),
    "sb_csr_nnz": (_i64, [_vp]),
    "sb_csr_num_rows": (_i64, [_vp]),
    "sb_csr_num_cols": (_i64, [_vp]),
    "sb_csr_pattern_dev": (_i, [_vp, _pp, _pp]),
    "sb_csr_pattern_to_host": (_i, [_vp, _vp, _vp]),
    "sb_csr_assemble": (_i, [_vp, _vp, _vp, _vp]),
    "sb_csr_accumulate": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "sb_functional_create": (_i, [_vp, _i, _pp, _pp]),
    "sb_functional_add_boundary_integral": (_i, [_vp, _vp, _vp, _vp, _pp]),
    "sb_functional_add_integral_on": (_i, [_vp, _vp, _vp, _vp, _pp]),
    "sb_functional_destroy": (_i, [_vp]),
    "sb_functional_add_integral": (_i, [_vp, _vp, _vp]),
    "sb_functional_evaluate": (_i, [_vp, _d, _i, _pp, _i, _vp, _i, _i, _vp]),
    "sb_functional_jvp": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp]),
    "sb_functional_wait": (_i, [_vp]),
    "sb_functional_assemble": (_i, [_vp, _i, _pp, _vp, _i, _vp]),
    "sb_halo_create": (_i, [_i64, _i64, _vp, _i, _vp, _vp, _vp, _vp, _pp]),
    "sb_halo_destroy": (_i, [_vp]),
    "sb_halo_buffers": (_i, [_vp, _pp, _pp]),
    "sb_halo_set_buffers": (_i, [_vp, _vp, _vp]),
    "sb_halo_buffer_doubles": (_i64, [_vp]),
    "sb_halo_enable_peer": (_i, [_vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "sb_halo_P_peer": (_i, [_vp, _vp, _vp, _vp]),
    "sb_halo_Pt_peer": (_i, [_vp, _vp, _vp, _vp]),
    "sb_halo_P_peer_put": (_i, [_vp, _vp, _vp]),
    "sb_halo_P_peer_finish": (_i, [_vp, _vp, _vp, _vp]),
    "sb_halo_Pt_peer_put": (_i, [_vp, _vp, _vp]),
    "sb_halo_Pt_peer_finish": (_i, [_vp, _vp, _vp, _vp]),
    "sb_halo_owned_first": (_i, [_vp]),
    "sb_halo_P_peer_finish_inplace": (_i, [_vp, _vp, _vp]),
    "sb_halo_P_peer_inplace_multi": (_i, [_i, _vp, _vp, _vp]),
    "sb_halo_Pt_peer_inplace": (_i, [_vp, _vp, _vp]),
    "sb_halo_P_peer_put_multi": (_i, [_i, _vp, _vp, _vp]),
    "sb_halo_P_peer_finish_inplace_multi": (_i, [_i, _vp, _vp, _vp]),
    "sb_functional_evaluate_begin": (_i, [_vp, _d, _i, _vp, _i, _i64, _vp]),
    "sb_functional_evaluate_end": (_i, [_vp, _vp, _vp]),
    "sb_halo_Pt_peer_finish_inplace": (_i, [_vp, _vp, _vp]),
    "sb_halo_peer_status": (_i, [_vp, _vp]),
    "sb_halo_P_pack": (_i, [_vp, _vp, _vp]),
    "sb_halo_P_finish": (_i, [_vp, _vp, _vp, _vp]),
    "sb_halo_Pt_pack": (_i, [_vp, _vp, _vp]),
    "sb_halo_Pt_finish": (_i, [_vp, _vp, _vp, _vp]),
    # include/serac_b200_comm.h
    "sb_comm_unique_id": (_i, [_vp]),
    "sb_comm_create": (_i, [_vp, _i, _i, _pp]),
    "sb_comm_destroy": (_i, [_vp]),
    "sb_comm_rank": (_i, [_vp]),
    "sb_comm_size": (_i, [_vp]),
    "sb_comm_allreduce_sum": (_i, [_vp, _vp, _i64, _vp]),
    "sb_comm_barrier": (_i, [_vp, _vp]),
    "sb_halo_set_neighbor_ranks": (_i, [_vp, _vp]),
    "sb_halo_P_nccl": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "sb_halo_Pt_nccl": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "sb_halo_enable_peer_ipc": (_i, [_vp, _vp]),
    "sb_par_csr_create": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _pp]),
    "sb_par_csr_destroy": (_i, [_vp]),
    "sb_par_csr_mult": (_i, [_vp, _vp, _vp, _vp]),
    "sb_par_csr_diagonal": (_i, [_vp, _vp, _vp]),
    "sb_par_dot": (_i, [_vp, _vp, _vp, _i64, C.POINTER(_d), _vp]),
    "sb_par_cg_solve": (_i, [_vp, _vp, _vp, _vp, _d, _d, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_d), _vp]),
    # include/serac_b200_solvers.h
    "sb_vector_set_subvector": (_i, [_vp, _vp, _i64, _d, _vp]),
    "sb_vector_copy_subvector": (_i, [_vp, _vp, _vp, _i64, _vp]),
    "sb_vector_axpy": (_i, [_vp, _d, _vp, _i64, _vp]),
    "sb_vector_dot": (_i, [_vp, _vp, _i64, C.POINTER(_d), _vp]),
    "sb_csr_spmv": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "sb_csr_diagonal": (_i, [_vp, _vp, _vp, _vp]),
    "sb_csr_row_l1_norms": (_i, [_vp, _vp, _vp, _vp]),
    "sb_csr_eliminate_rows_cols": (_i, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "sb_csr_eliminate_bc": (_i, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "sb_functional_gradient_operator_create": (_i, [_vp, _i, _i64, _vp, _i64, _pp]),
    "sb_functional_gradient_operator_destroy": (_i, [_vp]),
    "sb_functional_gradient_operator_apply": (_i, [_vp, _vp, _vp, _vp]),
    "sb_cg_create": (_i, [_i64, _pp]),
    "sb_cg_destroy": (_i, [_vp]),
    "sb_cg_solve_csr": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sb_cg_solve_operator": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sb_jacobi_setup": (_i, [_vp, _vp, _i64, _vp]),
    "sb_gmres_create": (_i, [_i64, _i, _pp]),
    "sb_gmres_destroy": (_i, [_vp]),
    "sb_gmres_solve_csr": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sb_gmres_solve_operator": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
}


class CgOptions(C.Structure):
    _fields_ = [("rel_tol", _d), ("abs_tol", _d), ("max_iter", _i), ("iterative_mode", _i), ("check_every", _i)]


class CgResult(C.Structure):
    _fields_ = [("converged", _i), ("final_iter", _i), ("initial_norm", _d), ("final_norm", _d)]

_lib = None


class SeracB200Error(RuntimeError):
    pass


def lib():
    """Load libserac_b200.so; raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SeracB200Error(f"{LIB_PATH} is missing: build it with `make -C serac_b200/csrc` "
                                 "(there is no CPU fallback for the Functional path)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise SeracB200Error(f"serac_b200 error {rc}: {lib().sb_last_error().decode()}")


def _np_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _dev_ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def geometry_dimension(geom):
    return {SEGMENT: 1, TRIANGLE: 2, SQUARE: 2, TETRAHEDRON: 3, CUBE: 3}[geom]


def num_nodes(geom, p):
    """H1 nodes per element"""
    return {SEGMENT: p + 1, TRIANGLE: (p + 1) * (p + 2) // 2, SQUARE: (p + 1)**2,
            TETRAHEDRON: (p + 1) * (p + 2) * (p + 3) // 6, CUBE: (p + 1)**3}[geom]


def num_quadrature_points(geom, q):
    """geometry.hpp:25-43"""
    return {SEGMENT: q, TRIANGLE: q * (q + 1) // 2, SQUARE: q * q, TETRAHEDRON: q * (q + 1) * (q + 2) // 6,
            CUBE: q**3}[geom]


def lexicographic_permutation(geom, p):
    n = num_nodes(geom, p)
    out = np.empty(n, dtype=np.int32)
    check(lib().sb_h1_lexicographic_permutation(geom, p, _np_ptr(out)))
    return out


def h1_dof_info(geom, p, native_dofs):
    """GetElementRestriction for H1 (element_restriction.cpp:216-277): native mfem element dofs -> packed DoF words."""
    native = np.ascontiguousarray(native_dofs, dtype=np.int32)
    out = np.empty(native.shape, dtype=np.uint64)
    check(lib().sb_h1_lexicographic_dof_info(geom, p, _np_ptr(native), native.shape[0], _np_ptr(out)))
    return out


def l2_face_dof_info(face_geom, p, elem_geom, elem_dofs, offsets, side, orientation):
    """face restriction table of a discontinuous (L2) space, FaceType::BOUNDARY or INTERIOR (element_restriction.cpp:
    304-355): one row per (face, adjacent element) pair; host only"""
    eg, ed, so, si, orr = (np.ascontiguousarray(a, dtype=t) for a, t in
                           ((elem_geom, np.int32), (elem_dofs, np.int32), (offsets, np.int64), (side, np.int32),
                            (orientation, np.int32)))
    n = len(si)
    out = np.empty((n, num_nodes(face_geom, p)), dtype=np.uint64)
    check(lib().sb_l2_face_dof_info(int(face_geom), int(p), n, _np_ptr(eg), _np_ptr(ed), _np_ptr(so), _np_ptr(si),
                                    _np_ptr(orr), _np_ptr(out)))
    return out


class ElementRestriction:
    """Mirror of serac::ElementRestriction (element_restriction.hpp:150-215) for one geometry."""

    def __init__(self, dof_info, components, num_nodes, ordering=BYVDIM):
        info = np.ascontiguousarray(dof_info, dtype=np.uint64)
        self.num_elements, self.nodes_per_elem = info.shape
        self.components, self.num_nodes, self.ordering = components, int(num_nodes), ordering
        h = C.c_void_p()
        check(lib().sb_restriction_create(_np_ptr(info), self.num_elements, self.nodes_per_elem, components,
                                          self.num_nodes, ordering, C.byref(h)))
        self.h = h
        self.node_stride = components if ordering == BYVDIM else 1
        self.comp_stride = 1 if ordering == BYVDIM else self.num_nodes
        t = C.c_void_p()
        check(lib().sb_restriction_table_dev(self.h, C.byref(t)))
        self.table_dev = t.value

    def __del__(self):
        try:
            lib().sb_restriction_destroy(self.h)
        except Exception:
            pass

    def ESize(self):
        return lib().sb_restriction_esize(self.h)

    def LSize(self):
        return lib().sb_restriction_lsize(self.h)

    def Gather(self, L, E=None):
        import torch
        if E is None:
            E = torch.empty(self.ESize(), dtype=torch.float64, device=L.device)
        check(lib().sb_restriction_gather(self.h, _dev_ptr(L), _dev_ptr(E), _stream_ptr()))
        return E

    def ScatterAdd(self, E, L, node_major=False):
        """node_major: E is [element][node][comp] (the Functional's internal layout) instead of [element][comp][node]"""
        if node_major:
            check(lib().sb_restriction_scatter_layout(self.h, _dev_ptr(E), _dev_ptr(L), 1, 1, _stream_ptr()))
        else:
            check(lib().sb_restriction_scatter_add(self.h, _dev_ptr(E), _dev_ptr(L), _stream_ptr()))
        return L

    def Scatter(self, E, L, node_major=False):
        if node_major:
            check(lib().sb_restriction_scatter_layout(self.h, _dev_ptr(E), _dev_ptr(L), 0, 1, _stream_ptr()))
        else:
            check(lib().sb_restriction_scatter(self.h, _dev_ptr(E), _dev_ptr(L), _stream_ptr()))
        return L

    def field(self, L):
        return FieldArg(L.data_ptr(), self.table_dev, self.node_stride, self.comp_stride)


def geometric_factors(geom, mesh_order, q, mesh_restriction, nodes_L, num_elements=None, elements=None):
    """GeometricFactors (geometric_factors.cpp:62-138) -> (X_q [ne][dim][nq], J_q [ne][dim][dim][nq]) on device."""
    import torch
    dim = geometry_dimension(geom)
    sdim = mesh_restriction.components  # == dim for domain elements, dim + 1 for the faces of a boundary integral
    ne = mesh_restriction.num_elements if num_elements is None else num_elements
    nq = num_quadrature_points(geom, q)
    X = torch.empty((ne, sdim, nq), dtype=torch.float64, device=nodes_L.device)
    J = torch.empty((ne, dim, sdim, nq), dtype=torch.float64, device=nodes_L.device)
    check(lib().sb_geometric_factors(geom, mesh_order, q, mesh_restriction.h, _dev_ptr(nodes_L),
                                     _dev_ptr(elements), ne, _dev_ptr(X), _dev_ptr(J), _stream_ptr()))
    return X, J


def named_integrands():
    L = lib()
    return [L.sb_named_integrand_name(i).decode() for i in range(L.sb_named_integrand_count())]


class Integral:
    """Mirror of serac::Integral (integral.hpp:24-168) created from a named (compiled-in) integrand."""

    def __init__(self, name, params, X_q, J_q, num_elements, elements=None):
        L = lib()
        tp = C.POINTER(KernelTable)()
        check(L.sb_named_kernel_table(name.encode(), C.byref(tp)))
        self.table = tp.contents
        params = np.ascontiguousarray(params, dtype=np.float64)
        blob = (C.c_ubyte * self.table.qf_bytes)()
        check(L.sb_named_qf_blob(name.encode(), _np_ptr(params), len(params), blob, self.table.qf_bytes))
        el = None if elements is None else np.ascontiguousarray(elements, dtype=np.int32)
        h = C.c_void_p()
        check(L.sb_integral_create(tp, blob, None if el is None else _np_ptr(el), num_elements, _dev_ptr(X_q),
                                   _dev_ptr(J_q), C.byref(h)))
        self.h = h
        self._keep = (X_q, J_q)
        self.num_elements = num_elements
        self.nq = num_quadrature_points(self.table.geom, self.table.q)
        self.qdata = None

    def __del__(self):
        try:
            lib().sb_integral_destroy(self.h)
        except Exception:
            pass

    def set_qdata(self, qdata):
        self.qdata = qdata
        check(lib().sb_integral_set_qdata(self.h, _dev_ptr(qdata)))

    def Mult(self, t, fields, r_E, wrt=NO_DIFFERENTIATION, update_state=False, node_major=False):
        arr = (FieldArg * len(fields))(*fields)
        check(lib().sb_integral_evaluate_layout(self.h, float(t), arr, _dev_ptr(r_E), wrt, int(update_state),
                                                int(node_major), _stream_ptr()))
        return r_E

    def GradientMult(self, wrt, du_field, dr_E, node_major=False):
        check(lib().sb_integral_jvp_layout(self.h, wrt, C.byref(du_field), _dev_ptr(dr_E), int(node_major),
                                           _stream_ptr()))
        return dr_E

    def ComputeElementGradients(self, wrt, K_e):
        check(lib().sb_integral_element_gradient(self.h, wrt, _dev_ptr(K_e), _stream_ptr()))
        return K_e

    def derivative_entries(self, wrt):
        return lib().sb_integral_derivative_entries(self.h, wrt)

    def derivative_stored_entries(self, wrt):
        """doubles per quadrature point actually stored (symmetric tangents: the upper triangle only)"""
        return lib().sb_integral_derivative_stored_entries(self.h, wrt)

    def derivative_mask(self, wrt):
        """bits 0-3: blocks d00, d01, d10, d11 present; bit 4 (16): d11 stored as the upper triangle"""
        return lib().sb_integral_derivative_mask(self.h, wrt)

    def export_derivatives(self, wrt):
        import torch
        out = torch.empty((self.num_elements, self.nq, self.derivative_entries(wrt)), dtype=torch.float64,
                          device="cuda")
        check(lib().sb_integral_export_derivatives(self.h, wrt, _dev_ptr(out), _stream_ptr()))
        return out


class CsrPattern:
    """Mirror of GradientAssemblyLookupTables (dof_numbering.hpp:253-380)."""

    def __init__(self, test, trial, handle=None):
        """test / trial: one ElementRestriction each, or -- meshes of mixed element geometry -- equally long lists with
        one restriction per geometry (BlockElementRestriction, element_restriction.cpp:471-536)"""
        if handle is None:
            handle = C.c_void_p()
            if isinstance(test, (list, tuple)):
                ta = (C.c_void_p * len(test))(*[r.h for r in test])
                ka = (C.c_void_p * len(trial))(*[r.h for r in trial])
                check(lib().sb_csr_create_blocks(len(test), ta, ka, C.byref(handle)))
            else:
                check(lib().sb_csr_create(test.h, trial.h, C.byref(handle)))
        self.h = handle
        self._keep = (test, trial)  # the pattern refers to the restrictions it was built from
        self.nnz = lib().sb_csr_nnz(self.h)
        self.num_rows = lib().sb_csr_num_rows(self.h)

    def __del__(self):
        try:
            lib().sb_csr_destroy(self.h)
        except Exception:
            pass

    def transpose(self):
        """the pattern of A^T (HypreParMatrix::Transpose, solid_mechanics.hpp:1365): a CsrPattern whose
        transpose_values(values) reorders the values of A into those of A^T"""
        h = C.c_void_p()
        check(lib().sb_csr_transpose(self.h, C.byref(h)))
        t = CsrPattern(self._keep[1], self._keep[0], handle=h)
        t._of = self
        return t

    def transpose_values(self, values, out=None):
        import torch
        if out is None:
            out = torch.empty_like(values)
        check(lib().sb_csr_transpose_values(self.h, _dev_ptr(values), _dev_ptr(out), _stream_ptr()))
        return out

    def to_host(self):
        rp = np.empty(self.num_rows + 1, dtype=np.int32)
        ci = np.empty(self.nnz, dtype=np.int32)
        check(lib().sb_csr_pattern_to_host(self.h, _np_ptr(rp), _np_ptr(ci)))
        return rp, ci

    def assemble(self, K_e, values=None):
        import torch
        if values is None:
            values = torch.empty(self.nnz, dtype=torch.float64, device=K_e.device)
        check(lib().sb_csr_assemble(self.h, _dev_ptr(K_e), _dev_ptr(values), _stream_ptr()))
        return values


def csr_pattern_host(test_table, test_components, test_num_nodes, trial_table, trial_components, trial_num_nodes,
                     test_ordering=BYVDIM, trial_ordering=BYVDIM):
    """row_ptr / col_ind of GradientAssemblyLookupTables from raw restriction tables, computed on the host by the same
    code sb_csr_create runs before uploading (no device needed)."""
    tt = np.ascontiguousarray(test_table, dtype=np.int32)
    kt = np.ascontiguousarray(trial_table, dtype=np.int32)
    ne = tt.shape[0]
    rp = np.zeros(test_num_nodes * test_components + 1, dtype=np.int32)
    nnz = _i64(0)
    args = [_np_ptr(tt), tt.shape[1], test_components, test_num_nodes, test_ordering, _np_ptr(kt), kt.shape[1],
            trial_components, trial_num_nodes, trial_ordering, ne]
    check(lib().sb_csr_pattern_host(*args, _np_ptr(rp), None, C.byref(nnz)))
    ci = np.empty(nnz.value, dtype=np.int32)
    check(lib().sb_csr_pattern_host(*args, _np_ptr(rp), _np_ptr(ci), C.byref(nnz)))
    return rp, ci


def debug_assembly_plan_host(test_table, test_components, test_num_nodes, trial_table, trial_components,
                             trial_num_nodes, K_e, row_ptr, nnz, elements=None, acc_cap=1600, trial_blocks=1,
                             test_ordering=BYVDIM, trial_ordering=BYVDIM):
    """test hook: walk the fused assembly's index tables (groups, units, row positions) over element matrices given by
    the caller, on the host.  K_e: [domain position][trial vdof][test vdof].  -> (values, number of groups)"""
    tt = np.ascontiguousarray(test_table, dtype=np.int32)
    kt = np.ascontiguousarray(trial_table, dtype=np.int32)
    K_e = np.ascontiguousarray(K_e, dtype=np.float64)
    rp = np.ascontiguousarray(row_ptr, dtype=np.int32)
    el = None if elements is None else np.ascontiguousarray(elements, dtype=np.int32)
    nd = tt.shape[0] if el is None else len(el)
    values = np.zeros(nnz)
    ng = _i(0)
    check(lib().sb_debug_assembly_plan_host(_np_ptr(tt), tt.shape[1], test_components, test_num_nodes, test_ordering,
                                            _np_ptr(kt), kt.shape[1], trial_components, trial_num_nodes, trial_ordering,
                                            tt.shape[0], None if el is None else _np_ptr(el), nd, acc_cap, trial_blocks,
                                            _np_ptr(K_e),
                                            _np_ptr(rp), _np_ptr(values), C.byref(ng)))
    return values, ng.value


class Functional:
    """Mirror of serac::Functional's evaluation orchestration at one rank (functional.hpp:406-513,569-652)."""

    def __init__(self, test, trials):
        self.test, self.trials = test, list(trials)
        arr = (C.c_void_p * len(trials))(*[t.h for t in trials])
        h = C.c_void_p()
        check(lib().sb_functional_create(test.h, len(trials), arr, C.byref(h)))
        self.h = h
        self.integrals = []
        self._csr = {}

    def __del__(self):
        try:
            lib().sb_functional_destroy(self.h)
        except Exception:
            pass

    def AddDomainIntegral(self, integral, depends_on):
        idx = np.ascontiguousarray(depends_on, dtype=np.int32)
        check(lib().sb_functional_add_integral(self.h, integral.h, _np_ptr(idx)))
        self.integrals.append(integral)

    def AddBoundaryIntegral(self, integral, depends_on, test_faces, trial_faces):
        """functional.hpp:328-354: test_faces / trial_faces[i] are the boundary-face ElementRestrictions of the test space
        and of the space feeding the integral's i-th trial"""
        idx = np.ascontiguousarray(depends_on, dtype=np.int32)
        arr = (C.c_void_p * len(trial_faces))(*[r.h for r in trial_faces])
        check(lib().sb_functional_add_boundary_integral(self.h, integral.h, _np_ptr(idx), test_faces.h, arr))
        self.integrals.append(integral)
        self._keep = getattr(self, "_keep", []) + [test_faces, list(trial_faces)]

    def AddDomainIntegralOn(self, integral, depends_on, test_block, trial_blocks):
        """an integral on the elements of ONE geometry of a mixed mesh (integral.hpp:254-257): test_block / trial_blocks[i]
        are the restrictions of the spaces on those elements; assemble() then builds its pattern over every block"""
        idx = np.ascontiguousarray(depends_on, dtype=np.int32)
        arr = (C.c_void_p * len(trial_blocks))(*[r.h for r in trial_blocks])
        check(lib().sb_functional_add_integral_on(self.h, integral.h, _np_ptr(idx), test_block.h, arr))
        self.integrals.append(integral)
        self._keep = getattr(self, "_keep", []) + [test_block, list(trial_blocks)]
        self._blocks = getattr(self, "_blocks", []) + [(test_block, list(depends_on), list(trial_blocks))]

    def wait(self):
        """results of asynchronous host-buffer calls are valid after this"""
        check(lib().sb_functional_wait(self.h))

    def __call__(self, t, inputs, wrt=NO_DIFFERENTIATION, out=None, update_state=False, async_host=False):
        """inputs: list of torch CUDA tensors (device path) or numpy arrays (host path, copies inside the call).
        async_host: the result's device -> host copy overlaps later calls; call wait() before reading `out`."""
        import torch
        on_host = isinstance(inputs[0], np.ndarray)
        n = self.test.LSize()
        if on_host:
            ins = [np.ascontiguousarray(a, dtype=np.float64) for a in inputs]
            ptrs = (C.c_void_p * len(ins))(*[a.ctypes.data for a in ins])
            if out is None:
                out = np.empty(n)
            mode = 2 if async_host else 1
            check(lib().sb_functional_evaluate(self.h, float(t), wrt, ptrs, mode, _np_ptr(out), mode, int(update_state),
                                               _stream_ptr()))
        else:
            ptrs = (C.c_void_p * len(inputs))(*[a.data_ptr() for a in inputs])
            if out is None:
                out = torch.empty(n, dtype=torch.float64, device=inputs[0].device)
            check(lib().sb_functional_evaluate(self.h, float(t), wrt, ptrs, 0, _dev_ptr(out), 0, int(update_state),
                                               _stream_ptr()))
        return out

    def evaluate_begin(self, t, inputs, wrt=NO_DIFFERENTIATION, update_state=False, split_element=0):
        """first half of __call__ (device tensors): the elements [0, split_element) of the whole-mesh domain integral"""
        ptrs = (C.c_void_p * len(inputs))(*[a.data_ptr() for a in inputs])
        self._pending_inputs = list(inputs)  # keep alive until evaluate_end
        check(lib().sb_functional_evaluate_begin(self.h, float(t), wrt, ptrs, int(update_state), int(split_element),
                                                 _stream_ptr()))

    def evaluate_end(self, out=None):
        import torch
        if out is None:
            out = torch.empty(self.test.LSize(), dtype=torch.float64, device=self._pending_inputs[0].device)
        check(lib().sb_functional_evaluate_end(self.h, _dev_ptr(out), _stream_ptr()))
        self._pending_inputs = None
        return out

    def GradientMult(self, wrt, du, out=None, async_host=False):
        import torch
        n = self.test.LSize()
        if isinstance(du, np.ndarray):
            du = np.ascontiguousarray(du, dtype=np.float64)
            if out is None:
                out = np.empty(n)
            mode = 2 if async_host else 1
            check(lib().sb_functional_jvp(self.h, wrt, _np_ptr(du), mode, _np_ptr(out), mode, _stream_ptr()))
        else:
            if out is None:
                out = torch.empty(n, dtype=torch.float64, device=du.device)
            check(lib().sb_functional_jvp(self.h, wrt, _dev_ptr(du), 0, _dev_ptr(out), 0, _stream_ptr()))
        return out

    def assemble(self, wrt, on_host=False):
        """-> (CsrPattern, values): the local CSR triple the reference wraps in mfem::SparseMatrix (functional.hpp:649)."""
        import torch
        if wrt not in self._csr:
            blocks = [(t, ks[dep.index(wrt)]) for (t, dep, ks) in getattr(self, "_blocks", []) if wrt in dep]
            if blocks:  # mixed geometries: the pattern is the union over the per-geometry blocks
                has_plain = len(self.integrals) > len(getattr(self, "_blocks", []))
                tests = ([self.test] if has_plain else []) + [b[0] for b in blocks]
                trials = ([self.trials[wrt]] if has_plain else []) + [b[1] for b in blocks]
                self._csr[wrt] = CsrPattern(tests, trials)
            else:
                self._csr[wrt] = CsrPattern(self.test, self.trials[wrt])
        pat = self._csr[wrt]
        hp = C.c_void_p(pat.h.value)
        if on_host:
            vals = np.empty(pat.nnz)
            check(lib().sb_functional_assemble(self.h, wrt, C.byref(hp), _np_ptr(vals), 1, _stream_ptr()))
        else:
            vals = torch.empty(pat.nnz, dtype=torch.float64, device="cuda")
            check(lib().sb_functional_assemble(self.h, wrt, C.byref(hp), _dev_ptr(vals), 0, _stream_ptr()))
        return pat, vals


# ---- the Newton step's linear algebra on the device (include/serac_b200_solvers.h) ------------------------------------
def _ess_tensor(ess, device="cuda"):
    import torch
    if isinstance(ess, torch.Tensor):
        return ess.to(device=device, dtype=torch.int32).contiguous()
    return torch.as_tensor(np.ascontiguousarray(ess, dtype=np.int32), device=device)


def set_subvector(v, ess, value=0.0):
    """r.SetSubVector(bcs_.allEssentialTrueDofs(), 0.0)  (solid_mechanics.hpp:1129)"""
    ess = _ess_tensor(ess, v.device)
    check(lib().sb_vector_set_subvector(_dev_ptr(v), _dev_ptr(ess), ess.numel(), float(value), _stream_ptr()))
    return v


def vector_axpy(y, a, x):
    """y += a x on the device (the Newton update)"""
    check(lib().sb_vector_axpy(_dev_ptr(y), float(a), _dev_ptr(x), y.numel(), _stream_ptr()))
    return y


def vector_dot(x, y):
    """(x, y) with the library's deterministic reduction"""
    out = C.c_double(0.0)
    check(lib().sb_vector_dot(_dev_ptr(x), _dev_ptr(y), x.numel(), C.byref(out), _stream_ptr()))
    return out.value


def copy_subvector(v, src, ess):
    """v[ess] = src[ess] (the essential rows of the warm-start right-hand side, solid_mechanics.hpp:1757-1760)"""
    ess = _ess_tensor(ess, v.device)
    check(lib().sb_vector_copy_subvector(_dev_ptr(v), _dev_ptr(src), _dev_ptr(ess), ess.numel(), _stream_ptr()))
    return v


def csr_spmv(pattern, values, x, out=None):
    import torch
    if out is None:
        out = torch.empty(pattern.num_rows, dtype=torch.float64, device=x.device)
    check(lib().sb_csr_spmv(pattern.h, _dev_ptr(values), _dev_ptr(x), _dev_ptr(out), _stream_ptr()))
    return out


def csr_diagonal(pattern, values):
    import torch
    d = torch.empty(pattern.num_rows, dtype=torch.float64, device=values.device)
    check(lib().sb_csr_diagonal(pattern.h, _dev_ptr(values), _dev_ptr(d), _stream_ptr()))
    return d


def csr_row_l1_norms(pattern, values):
    """sum_j |A(i,j)|: the scaling of HypreSmoother::l1Jacobi (Preconditioner::HypreL1Jacobi)"""
    import torch
    d = torch.empty(pattern.num_rows, dtype=torch.float64, device=values.device)
    check(lib().sb_csr_row_l1_norms(pattern.h, _dev_ptr(values), _dev_ptr(d), _stream_ptr()))
    return d


def jacobi_setup(diag):
    import torch
    dinv = torch.empty_like(diag)
    check(lib().sb_jacobi_setup(_dev_ptr(diag), _dev_ptr(dinv), diag.numel(), _stream_ptr()))
    return dinv


def eliminate_rows_cols(pattern, values, ess, want_Ae=True):
    """bcs_.eliminateAllEssentialDofsFromMatrix(*J_) (boundary_condition_manager.hpp:91-94): values is modified in place;
    returns Ae's values in the same pattern (or None)"""
    import torch
    ess = _ess_tensor(ess, values.device)
    Ae = torch.empty_like(values) if want_Ae else None
    check(lib().sb_csr_eliminate_rows_cols(pattern.h, _dev_ptr(values), _dev_ptr(ess), ess.numel(), _dev_ptr(Ae),
                                           _stream_ptr()))
    return Ae


def eliminate_bc(pattern, values, values_e, ess, X, B):
    """mfem::EliminateBC(A, Ae, ess, X, B) (solid_mechanics.hpp:1756): B is modified in place"""
    ess = _ess_tensor(ess, values.device)
    check(lib().sb_csr_eliminate_bc(pattern.h, _dev_ptr(values), _dev_ptr(values_e), _dev_ptr(ess), ess.numel(),
                                    _dev_ptr(X), _dev_ptr(B), _stream_ptr()))
    return B


class GradientOperator:
    """Matrix-free [J_ii 0; 0 I]: Gradient::Mult (functional.hpp:556-559) with essential dofs eliminated."""

    def __init__(self, functional, wrt, ess=()):
        self.F, self.wrt = functional, wrt
        self.n = functional.trials[wrt].LSize()
        self._ess = _ess_tensor(ess)
        h = C.c_void_p()
        check(lib().sb_functional_gradient_operator_create(functional.h, wrt, self.n, _dev_ptr(self._ess),
                                                           self._ess.numel(), C.byref(h)))
        self.h = h

    def __del__(self):
        try:
            lib().sb_functional_gradient_operator_destroy(self.h)
        except Exception:
            pass

    def Mult(self, x, out=None):
        import torch
        if out is None:
            out = torch.empty(self.n, dtype=torch.float64, device=x.device)
        check(lib().sb_functional_gradient_operator_apply(self.h, _dev_ptr(x), _dev_ptr(out), _stream_ptr()))
        return out


class CGSolver:
    """Mirror of mfem::CGSolver (mfem/linalg/solvers.cpp:719-897) as serac::buildLinearSolver configures it."""

    def __init__(self, n, rel_tol=1e-12, abs_tol=0.0, max_iter=1000, check_every=1):
        self.n = n
        self.opt = CgOptions(rel_tol, abs_tol, max_iter, 0, check_every)
        h = C.c_void_p()
        check(lib().sb_cg_create(n, C.byref(h)))
        self.h = h
        self.result = CgResult()

    def __del__(self):
        try:
            lib().sb_cg_destroy(self.h)
        except Exception:
            pass

    def _finish(self):
        r = self.result
        return dict(converged=bool(r.converged), final_iter=r.final_iter, initial_norm=r.initial_norm,
                    final_norm=r.final_norm)

    def Mult(self, operator, b, x=None, dinv=None, values=None, iterative_mode=False):
        """operator: a CsrPattern (with `values`) or a GradientOperator.  -> x, info"""
        import torch
        if x is None:
            x = torch.zeros(self.n, dtype=torch.float64, device=b.device)
        self.opt.iterative_mode = int(iterative_mode)
        if isinstance(operator, CsrPattern):
            check(lib().sb_cg_solve_csr(self.h, operator.h, _dev_ptr(values), _dev_ptr(dinv), _dev_ptr(b), _dev_ptr(x),
                                        C.byref(self.opt), C.byref(self.result), _stream_ptr()))
        else:
            apply = C.cast(lib().sb_functional_gradient_operator_apply, C.c_void_p)
            check(lib().sb_cg_solve_operator(self.h, apply, operator.h, _dev_ptr(dinv), _dev_ptr(b), _dev_ptr(x),
                                             C.byref(self.opt), C.byref(self.result), _stream_ptr()))
        return x, self._finish()


class GMRESSolver:
    """Mirror of mfem::GMRESSolver (mfem/linalg/solvers.cpp:981-1165): serac's default linear solver
    (solid_mechanics.hpp:50-57, heat_transfer.hpp:34-38), restart length kdim (mfem default 50)."""

    def __init__(self, n, kdim=50, rel_tol=1e-12, abs_tol=0.0, max_iter=1000):
        self.n = n
        self.opt = CgOptions(rel_tol, abs_tol, max_iter, 0, 1)
        h = C.c_void_p()
        check(lib().sb_gmres_create(n, kdim, C.byref(h)))
        self.h = h
        self.result = CgResult()

    def __del__(self):
        try:
            lib().sb_gmres_destroy(self.h)
        except Exception:
            pass

    def Mult(self, operator, b, x=None, dinv=None, values=None, iterative_mode=False):
        """operator: a CsrPattern (with `values`) or a GradientOperator.  -> x, info"""
        import torch
        if x is None:
            x = torch.zeros(self.n, dtype=torch.float64, device=b.device)
        self.opt.iterative_mode = int(iterative_mode)
        if isinstance(operator, CsrPattern):
            check(lib().sb_gmres_solve_csr(self.h, operator.h, _dev_ptr(values), _dev_ptr(dinv), _dev_ptr(b), _dev_ptr(x),
                                           C.byref(self.opt), C.byref(self.result), _stream_ptr()))
        else:
            apply = C.cast(lib().sb_functional_gradient_operator_apply, C.c_void_p)
            check(lib().sb_gmres_solve_operator(self.h, apply, operator.h, _dev_ptr(dinv), _dev_ptr(b), _dev_ptr(x),
                                                C.byref(self.opt), C.byref(self.result), _stream_ptr()))
        r = self.result
        return x, dict(converged=bool(r.converged), final_iter=r.final_iter, initial_norm=r.initial_norm,
                       final_norm=r.final_norm)
