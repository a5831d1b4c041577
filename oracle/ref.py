"""TEST INFRASTRUCTURE ONLY (oracle/): ctypes binding of oracle/_ref/libserac_ref.so -- the REFERENCE's own Functional
building blocks compiled from /root/reference (see oracle/ref_driver.hpp).  Used by tests (to validate the numpy
restatement and to generate golden fixtures) and by bench.py's cpu_baseline / --impl reference legs.  Never imported by
the product package."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libserac_ref.so")

GEOM = {"segment": 1, "triangle": 2, "quad": 3, "tet": 4, "hex": 5}
BYNODES, BYVDIM = 0, 1


def available():
    return os.path.exists(_LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_LIB_PATH, mode=os.RTLD_NOW | os.RTLD_DEEPBIND)  # python exports symbols that clash with mfem otherwise
        vp, i, d, l = C.c_void_p, C.c_int, C.c_double, C.c_long
        sig = {
            "ref_mesh_from_file": (vp, [C.c_char_p, i]),
            "ref_mesh_cartesian3d": (vp, [i, i, i, d, d, d]),
            "ref_mesh_cartesian2d": (vp, [i, i, d, d]),
            "ref_mesh_cartesian3d_tet": (vp, [i, i, i, d, d, d]),
            "ref_mesh_cartesian2d_tri": (vp, [i, i, d, d]),
            "ref_mesh_geom": (i, [vp]),
            "ref_mesh_destroy": (None, [vp]),
            "ref_mesh_dim": (i, [vp]),
            "ref_mesh_num_elements": (i, [vp]),
            "ref_mesh_perturb": (None, [vp, d]),
            "ref_mesh_nodes_order": (i, [vp]),
            "ref_mesh_nodes_ordering": (i, [vp]),
            "ref_mesh_nodes_size": (l, [vp]),
            "ref_mesh_nodes": (None, [vp, vp]),
            "ref_mesh_nodes_dof_info": (l, [vp, i, vp]),
            "ref_space_create": (vp, [vp, i, i, i]),
            "ref_space_create_l2": (vp, [vp, i, i, i]),
            "ref_space_destroy": (None, [vp]),
            "ref_space_vsize": (l, [vp]),
            "ref_space_num_nodes": (l, [vp]),
            "ref_space_esize": (l, [vp, i]),
            "ref_space_native_dofs": (l, [vp, i, vp]),
            "ref_space_dof_info": (l, [vp, i, vp]),
            "ref_space_gather": (None, [vp, i, vp, vp]),
            "ref_space_scatter_add": (None, [vp, i, vp, vp]),
            "ref_lex_perm": (i, [i, i, vp]),
            "ref_space_face_native_dofs": (l, [vp, i, vp]),
            "ref_space_face_dof_info": (l, [vp, i, vp]),
            "ref_space_face_dof_info_typed": (l, [vp, i, i, vp]),
            "ref_space_l2_face_inputs": (l, [vp, i, i, vp, vp, vp, vp, vp]),
            "ref_space_l2_face_inputs_size": (l, [vp, i, i]),
            "ref_space_face_esize": (l, [vp, i]),
            "ref_space_face_gather": (None, [vp, i, vp, vp]),
            "ref_space_face_scatter_add": (None, [vp, i, vp, vp]),
            "ref_mesh_nodes_face_dof_info": (l, [vp, i, vp]),
            "ref_mesh_nodes_face_native_dofs": (l, [vp, i, vp]),
            "ref_boundary_geometric_factors": (l, [vp, i, i, vp, vp]),
            "ref_csr_assemble_faces": (None, [vp, vp, vp, i, vp, vp]),
            "ref_geometric_factors": (None, [vp, i, i, vp, vp]),
            "ref_csr_create": (vp, [vp, vp]),
            "ref_csr_destroy": (None, [vp]),
            "ref_csr_nnz": (l, [vp]),
            "ref_csr_num_rows": (l, [vp]),
            "ref_csr_pattern": (None, [vp, vp, vp]),
            "ref_csr_assemble": (None, [vp, vp, vp, i, vp, vp]),
            "ref_eliminate_rows_cols": (None, [i, vp, vp, vp, vp, i, vp]),
            "ref_eliminate_bc": (None, [i, vp, vp, vp, vp, vp, i, vp, vp]),
            "ref_cg": (None, [i, vp, vp, vp, vp, vp, i, d, d, i, i, vp, vp, vp, vp]),
            "ref_spmv": (None, [i, vp, vp, vp, vp, vp]),
            "ref_gmres": (None, [i, vp, vp, vp, vp, vp, i, i, d, d, i, i, vp, vp, vp, vp]),
            "ref_num_kinds": (i, []),
            "ref_kind_name": (C.c_char_p, [i]),
            "ref_integral_create": (vp, [C.c_char_p, vp, vp, i]),
            "ref_integral_create_subset": (vp, [C.c_char_p, vp, vp, i, vp, i]),
            "ref_integral_destroy": (None, [vp]),
            "ref_integral_state_bytes": (C.c_size_t, [vp]),
            "ref_integral_qdata_count": (C.c_size_t, [vp]),
            "ref_integral_get_qdata": (None, [vp, vp]),
            "ref_integral_set_qdata": (None, [vp, vp]),
            "ref_integral_mult": (i, [vp, i, d, vp, i, vp, i, i]),
            "ref_integral_jvp": (i, [vp, i, i, vp, vp]),
            "ref_integral_element_gradient": (i, [vp, i, i, vp, l, l, l]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def num_quadrature_points(geom, q):
    """F/geometry.hpp:25-43"""
    return {1: q, 2: q * (q + 1) // 2, 3: q * q, 4: q * (q + 1) * (q + 2) // 6, 5: q**3}[geom]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Mesh:
    def __init__(self, handle):
        self.h = handle
        self.dim = lib().ref_mesh_dim(handle)
        self.ne = lib().ref_mesh_num_elements(handle)
        self.geom = lib().ref_mesh_geom(handle)  # mfem::Geometry::Type of the (single-geometry) mesh
        if self.geom < 0:
            raise ValueError("mixed-geometry meshes are not supported by this driver")

    @staticmethod
    def from_file(path, refinements=0):
        return Mesh(lib().ref_mesh_from_file(path.encode(), refinements))

    @staticmethod
    def cartesian3d(nx, ny, nz, sx=1.0, sy=1.0, sz=1.0):
        return Mesh(lib().ref_mesh_cartesian3d(nx, ny, nz, sx, sy, sz))

    @staticmethod
    def cartesian2d(nx, ny, sx=1.0, sy=1.0):
        return Mesh(lib().ref_mesh_cartesian2d(nx, ny, sx, sy))

    @staticmethod
    def cartesian3d_tet(nx, ny, nz, sx=1.0, sy=1.0, sz=1.0):
        return Mesh(lib().ref_mesh_cartesian3d_tet(nx, ny, nz, sx, sy, sz))

    @staticmethod
    def cartesian2d_tri(nx, ny, sx=1.0, sy=1.0):
        return Mesh(lib().ref_mesh_cartesian2d_tri(nx, ny, sx, sy))

    def perturb(self, amp):
        lib().ref_mesh_perturb(self.h, amp)

    def nodes(self):
        """(order, ordering, L-vector of nodal coordinates, dof_info table of the nodes space)"""
        L = lib()
        n = L.ref_mesh_nodes_size(self.h)
        x = np.empty(n)
        L.ref_mesh_nodes(self.h, _p(x))
        cnt = L.ref_mesh_nodes_dof_info(self.h, self.geom, None)
        info = np.empty(cnt, dtype=np.uint64)
        L.ref_mesh_nodes_dof_info(self.h, self.geom, _p(info))
        return L.ref_mesh_nodes_order(self.h), L.ref_mesh_nodes_ordering(self.h), x, info.reshape(self.ne, -1)

    @property
    def face_geom(self):
        return {GEOM["hex"]: GEOM["quad"], GEOM["tet"]: GEOM["triangle"]}.get(self.geom, GEOM["segment"])

    def num_boundary_faces(self):
        return lib().ref_boundary_geometric_factors(self.h, 1, self.face_geom, None, None)

    def boundary_geometric_factors(self, q):
        """GeometricFactors(EntireBoundary, q, face_geom, FaceType::BOUNDARY): X [nf][sdim][nq], J [nf][gdim][sdim][nq]"""
        nf, gd = self.num_boundary_faces(), self.dim - 1
        nq = num_quadrature_points(self.face_geom, q)
        X = np.empty((nf, self.dim, nq))
        J = np.empty((nf, gd, self.dim, nq))
        lib().ref_boundary_geometric_factors(self.h, q, self.face_geom, _p(X), _p(J))
        return X, J

    def nodes_face_tables(self):
        """(native face dofs, dof_info) of the mesh-nodes space on the boundary faces"""
        L, g = lib(), self.face_geom
        n = L.ref_mesh_nodes_face_native_dofs(self.h, g, None)
        nat = np.empty(n, dtype=np.int32)
        L.ref_mesh_nodes_face_native_dofs(self.h, g, _p(nat))
        info = np.empty(L.ref_mesh_nodes_face_dof_info(self.h, g, None), dtype=np.uint64)
        L.ref_mesh_nodes_face_dof_info(self.h, g, _p(info))
        nf = self.num_boundary_faces()
        return nat.reshape(nf, -1), info.reshape(nf, -1)

    def geometric_factors(self, q):
        nq = num_quadrature_points(self.geom, q)
        X = np.empty((self.ne, self.dim, nq))
        J = np.empty((self.ne, self.dim, self.dim, nq))
        lib().ref_geometric_factors(self.h, q, self.geom, _p(X), _p(J))
        return X, J


class Space:
    def __init__(self, mesh, p, vdim, ordering=BYVDIM, family="H1"):
        self.mesh, self.p, self.vdim, self.ordering, self.family = mesh, p, vdim, ordering, family
        create = lib().ref_space_create if family == "H1" else lib().ref_space_create_l2
        self.h = create(mesh.h, p, vdim, ordering)
        self.vsize = lib().ref_space_vsize(self.h)
        self.num_nodes = lib().ref_space_num_nodes(self.h)
        self.geom = mesh.geom
        self.esize = lib().ref_space_esize(self.h, self.geom)

    def native_dofs(self):
        n = lib().ref_space_native_dofs(self.h, self.geom, None)
        out = np.empty(n, dtype=np.int32)
        lib().ref_space_native_dofs(self.h, self.geom, _p(out))
        return out.reshape(self.mesh.ne, -1)

    def dof_info(self):
        n = lib().ref_space_dof_info(self.h, self.geom, None)
        out = np.empty(n, dtype=np.uint64)
        lib().ref_space_dof_info(self.h, self.geom, _p(out))
        return out.reshape(self.mesh.ne, -1)

    def gather(self, Lvec):
        E = np.empty(self.esize)
        lib().ref_space_gather(self.h, self.geom, _p(np.ascontiguousarray(Lvec)), _p(E))
        return E

    def scatter_add(self, E, Lvec):
        lib().ref_space_scatter_add(self.h, self.geom, _p(np.ascontiguousarray(E)), _p(Lvec))
        return Lvec


def l2_face_tables(space, face_geom, interior):
    """DG face restrictions (element_restriction.cpp:304-355): the reference's dof_info [pairs][nodes per face] for
    FaceType::BOUNDARY (interior=False) or FaceType::INTERIOR, and the per-(face, element) inputs mfem provides:
    dict(elem_geom, elem_dofs, offsets, side, orientation)"""
    L, t = lib(), 1 if interior else 0
    n = L.ref_space_face_dof_info_typed(space.h, face_geom, t, None)
    info = np.empty(n, dtype=np.uint64)
    L.ref_space_face_dof_info_typed(space.h, face_geom, t, _p(info))
    pairs = L.ref_space_l2_face_inputs(space.h, face_geom, t, None, None, None, None, None)
    nd = L.ref_space_l2_face_inputs_size(space.h, face_geom, t)
    ins = dict(elem_geom=np.empty(pairs, dtype=np.int32), elem_dofs=np.empty(nd, dtype=np.int32),
               offsets=np.empty(pairs + 1, dtype=np.int64), side=np.empty(pairs, dtype=np.int32),
               orientation=np.empty(pairs, dtype=np.int32))
    L.ref_space_l2_face_inputs(space.h, face_geom, t, _p(ins["elem_geom"]), _p(ins["elem_dofs"]), _p(ins["offsets"]),
                               _p(ins["side"]), _p(ins["orientation"]))
    return info.reshape(pairs, -1) if pairs else info.reshape(0, 0), ins


class FaceSpace:
    """the FaceType::BOUNDARY ElementRestriction of a Space (element_restriction.cpp:406-419)"""

    def __init__(self, space):
        self.space, self.geom = space, space.mesh.face_geom
        self.nf = space.mesh.num_boundary_faces()
        self.esize = lib().ref_space_face_esize(space.h, self.geom)

    def native_dofs(self):
        n = lib().ref_space_face_native_dofs(self.space.h, self.geom, None)
        out = np.empty(n, dtype=np.int32)
        lib().ref_space_face_native_dofs(self.space.h, self.geom, _p(out))
        return out.reshape(self.nf, -1)

    def dof_info(self):
        n = lib().ref_space_face_dof_info(self.space.h, self.geom, None)
        out = np.empty(n, dtype=np.uint64)
        lib().ref_space_face_dof_info(self.space.h, self.geom, _p(out))
        return out.reshape(self.nf, -1)

    def gather(self, Lvec):
        E = np.empty(self.esize)
        lib().ref_space_face_gather(self.space.h, self.geom, _p(np.ascontiguousarray(Lvec)), _p(E))
        return E

    def scatter_add(self, E, Lvec):
        lib().ref_space_face_scatter_add(self.space.h, self.geom, _p(np.ascontiguousarray(E)), _p(Lvec))
        return Lvec


def lex_perm(geom, p):
    out = np.empty(64, dtype=np.int32)
    n = lib().ref_lex_perm(geom, p, _p(out))
    return out[:n].copy()


class Csr:
    def __init__(self, test, trial):
        self.test, self.trial = test, trial
        self.h = lib().ref_csr_create(test.h, trial.h)
        self.nnz = lib().ref_csr_nnz(self.h)
        self.num_rows = lib().ref_csr_num_rows(self.h)

    def pattern(self):
        rp = np.empty(self.num_rows + 1, dtype=np.int32)
        ci = np.empty(self.nnz, dtype=np.int32)
        lib().ref_csr_pattern(self.h, _p(rp), _p(ci))
        return rp, ci

    def assemble(self, K_e):
        vals = np.zeros(self.nnz)
        lib().ref_csr_assemble(self.h, self.test.h, self.trial.h, self.test.geom, _p(np.ascontiguousarray(K_e)),
                               _p(vals))
        return vals


    def assemble_faces(self, K_e, values):
        """+= boundary-face element matrices (functional.hpp:610-644, Domain::Type::BoundaryElements)"""
        lib().ref_csr_assemble_faces(self.h, self.test.h, self.trial.h, self.test.mesh.face_geom,
                                     _p(np.ascontiguousarray(K_e)), _p(values))
        return values


def kinds():
    return [lib().ref_kind_name(i).decode() for i in range(lib().ref_num_kinds())]


class Integral:
    """One reference Integral (integral.hpp:24) on the entire mesh; E-vector level API."""

    def __init__(self, kind, mesh, params, boundary=False, element_mask=None):
        """element_mask: bool per mesh element -> the integral's domain is Domain::ofElements of the selected ones"""
        params = np.ascontiguousarray(params, dtype=np.float64)
        if element_mask is None:
            self.h = lib().ref_integral_create(kind.encode(), mesh.h, _p(params), len(params))
        else:
            m = np.ascontiguousarray(element_mask, dtype=np.uint8)
            assert len(m) == mesh.ne and not boundary
            self.h = lib().ref_integral_create_subset(kind.encode(), mesh.h, _p(params), len(params), _p(m), len(m))
        if not self.h:
            raise KeyError(kind)
        self.mesh = mesh
        self.geom = mesh.face_geom if boundary else mesh.geom

    def mult(self, t, inputs_E, out_size, wrt=-1, update_state=False):
        ins = [np.ascontiguousarray(a, dtype=np.float64) for a in inputs_E]
        arr = (C.c_void_p * len(ins))(*[a.ctypes.data for a in ins])
        out = np.zeros(out_size)
        rc = lib().ref_integral_mult(self.h, self.geom, float(t), arr, len(ins), _p(out), wrt, int(update_state))
        assert rc == 0
        return out

    def jvp(self, wrt, du_E, out_size):
        out = np.zeros(out_size)
        rc = lib().ref_integral_jvp(self.h, self.geom, wrt, _p(np.ascontiguousarray(du_E)), _p(out))
        assert rc == 0
        return out

    def element_gradient(self, wrt, ne, ncol, nrow):
        K = np.zeros((ne, ncol, nrow))
        rc = lib().ref_integral_element_gradient(self.h, self.geom, wrt, _p(K), ne, ncol, nrow)
        assert rc == 0
        return K

    def state_bytes(self):
        return lib().ref_integral_state_bytes(self.h)

    def get_qdata(self):
        n = lib().ref_integral_state_bytes(self.h) * lib().ref_integral_qdata_count(self.h)
        out = np.empty(n // 8)
        lib().ref_integral_get_qdata(self.h, _p(out))
        return out

    def set_qdata(self, data):
        lib().ref_integral_set_qdata(self.h, _p(np.ascontiguousarray(data)))


# ---- the reference stack's linear algebra for a Newton step (oracle/ref_solvers.cpp: the reference's own mfem) ----
def _csr_args(row_ptr, col_ind):
    rp = np.ascontiguousarray(row_ptr, dtype=np.int32)
    ci = np.ascontiguousarray(col_ind, dtype=np.int32)
    return len(rp) - 1, rp, ci


def eliminate_rows_cols(row_ptr, col_ind, values, ess):
    """SparseMatrix::EliminateRowCol(rc, Ae, DIAG_ONE) for every rc in ess -> (A values, Ae values) in A's pattern"""
    n, rp, ci = _csr_args(row_ptr, col_ind)
    A = np.array(values, dtype=np.float64)
    Ae = np.zeros_like(A)
    e = np.ascontiguousarray(ess, dtype=np.int32)
    lib().ref_eliminate_rows_cols(n, _p(rp), _p(ci), _p(A), _p(e), len(e), _p(Ae))
    return A, Ae


def eliminate_bc(row_ptr, col_ind, values, values_e, ess, X, B):
    n, rp, ci = _csr_args(row_ptr, col_ind)
    A = np.ascontiguousarray(values, dtype=np.float64)
    Ae = np.ascontiguousarray(values_e, dtype=np.float64)
    e = np.ascontiguousarray(ess, dtype=np.int32)
    x = np.ascontiguousarray(X, dtype=np.float64)
    b = np.array(B, dtype=np.float64)
    lib().ref_eliminate_bc(n, _p(rp), _p(ci), _p(A), _p(Ae), _p(e), len(e), _p(x), _p(b))
    return b


def cg(row_ptr, col_ind, values, b, jacobi=True, rel_tol=1e-12, abs_tol=0.0, max_iter=1000, x0=None):
    """mfem::CGSolver (+ OperatorJacobiSmoother) on an mfem::SparseMatrix -> x, info dict"""
    n, rp, ci = _csr_args(row_ptr, col_ind)
    A = np.ascontiguousarray(values, dtype=np.float64)
    bb = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros(n) if x0 is None else np.array(x0, dtype=np.float64)
    conv, it = C.c_int(0), C.c_int(0)
    n0, n1 = C.c_double(0), C.c_double(0)
    lib().ref_cg(n, _p(rp), _p(ci), _p(A), _p(bb), _p(x), int(jacobi), rel_tol, abs_tol, int(max_iter),
                 int(x0 is not None), C.byref(conv), C.byref(it), C.byref(n0), C.byref(n1))
    return x, dict(converged=bool(conv.value), final_iter=it.value, initial_norm=n0.value, final_norm=n1.value)


def spmv(row_ptr, col_ind, values, x):
    n, rp, ci = _csr_args(row_ptr, col_ind)
    A = np.ascontiguousarray(values, dtype=np.float64)
    xx = np.ascontiguousarray(x, dtype=np.float64)
    y = np.zeros(n)
    lib().ref_spmv(n, _p(rp), _p(ci), _p(A), _p(xx), _p(y))
    return y


def gmres(row_ptr, col_ind, values, b, jacobi=True, kdim=50, rel_tol=1e-12, abs_tol=0.0, max_iter=1000, x0=None):
    """mfem::GMRESSolver (+ OperatorJacobiSmoother) on an mfem::SparseMatrix -> x, info dict"""
    n, rp, ci = _csr_args(row_ptr, col_ind)
    A = np.ascontiguousarray(values, dtype=np.float64)
    bb = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros(n) if x0 is None else np.array(x0, dtype=np.float64)
    conv, it = C.c_int(0), C.c_int(0)
    n0, n1 = C.c_double(0), C.c_double(0)
    lib().ref_gmres(n, _p(rp), _p(ci), _p(A), _p(bb), _p(x), int(jacobi), int(kdim), rel_tol, abs_tol, int(max_iter),
                    int(x0 is not None), C.byref(conv), C.byref(it), C.byref(n0), C.byref(n1))
    return x, dict(converged=bool(conv.value), final_iter=it.value, initial_norm=n0.value, final_norm=n1.value)
