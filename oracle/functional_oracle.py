"""TEST INFRASTRUCTURE ONLY (oracle/): numpy restatement of the reference's Functional evaluation path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product never does.  Every
function cites the reference code it restates (paths relative to /root/reference, F/ = src/serac/numerics/functional/).
The restatement is PINNED: tests/test_oracle.py checks it (a) against the golden fixtures under tests/golden/ that
oracle/make_golden.py wrote with the reference's own kernels compiled from /root/reference (oracle/_ref), and (b) live
against oracle/_ref when that library is present.  Parity status: pinned (residual, Jacobian action, element matrices,
CSR pattern and values, restriction tables, geometric factors: <= 1e-13 relative / bit-exact for integers).

Vectorised over elements and quadrature points; forward-mode AD is a small numpy dual class (value array + array of
partials), which is the reference's scheme (F/dual.hpp, F/tuple_tensor_dual_functions.hpp:124-163) with dense partials.
"""
import numpy as np

SEGMENT, TRIANGLE, SQUARE, TETRAHEDRON, CUBE = 1, 2, 3, 4, 5

# ----------------------------------------------------------------------------------------------------------------------
# quadrature + 1-D bases (F/polynomials.hpp:42-55, 271-283, 493-540)
# ----------------------------------------------------------------------------------------------------------------------
_GL_NODES = {
    1: [0.50000000000000000],
    2: [0.2113248654051871, 0.7886751345948129],
    3: [0.1127016653792583, 0.500000000000000, 0.887298334620742],
    4: [0.0694318442029737, 0.330009478207572, 0.669990521792428, 0.930568155797026],
}
_GL_WEIGHTS = {
    1: [1.000000000000000],
    2: [0.500000000000000, 0.500000000000000],
    3: [0.277777777777778, 0.444444444444444, 0.277777777777778],
    4: [0.173927422568727, 0.326072577431273, 0.326072577431273, 0.173927422568727],
}
_SQRT5 = 2.23606797749978981


def gauss_lobatto_interpolation(n, x):
    if n == 1:
        return [1.0]
    if n == 2:
        return [1.0 - x, x]
    if n == 3:
        return [(-1.0 + x) * (-1.0 + 2.0 * x), -4.0 * (-1.0 + x) * x, x * (-1.0 + 2.0 * x)]
    if n == 4:
        s = _SQRT5
        return [-(-1.0 + x) * (1.0 + 5.0 * (-1.0 + x) * x), -0.5 * s * (5.0 + s - 10.0 * x) * (-1.0 + x) * x,
                -0.5 * s * (-1.0 + x) * x * (-5.0 + s + 10.0 * x), x * (1.0 + 5.0 * (-1.0 + x) * x)]
    raise ValueError(n)


def gauss_lobatto_interpolation_derivative(n, x):
    if n == 1:
        return [0.0]
    if n == 2:
        return [-1.0, 1.0]
    if n == 3:
        return [-3.0 + 4.0 * x, 4.0 - 8.0 * x, -1.0 + 4.0 * x]
    if n == 4:
        s = _SQRT5
        return [-6.0 + 5.0 * (4.0 - 3.0 * x) * x, 2.5 * (1.0 + s + 2.0 * x * (-1.0 - 3.0 * s + 3.0 * s * x)),
                -2.5 * (-1.0 + s + 2.0 * x * (1.0 - 3.0 * s + 3.0 * s * x)), 1.0 + 5.0 * x * (-2.0 + 3.0 * x)]
    raise ValueError(n)


def basis_tables(p, q):
    """B(qi, d), G(qi, d) of F/detail/hexahedron_H1.inl:96-130 (no weights) and the quadrature weights."""
    n = p + 1
    B = np.array([gauss_lobatto_interpolation(n, x) for x in _GL_NODES[q]])
    G = np.array([gauss_lobatto_interpolation_derivative(n, x) for x in _GL_NODES[q]])
    return B, G, np.array(_GL_WEIGHTS[q])


# ----------------------------------------------------------------------------------------------------------------------
# simplex rules and nodal bases (F/polynomials.hpp:57-260, 285-365; F/detail/triangle_H1.inl, tetrahedron_H1.inl)
# ----------------------------------------------------------------------------------------------------------------------
def _tet_rule(q):
    if q == 1:
        return [[0.25, 0.25, 0.25]], [0.166666666666666656]
    if q == 2:
        a, b = 0.585410196624967936, 0.138196601125011008
        return [[a, b, b], [b, a, b], [b, b, a], [b, b, b]], [0.0416666666666666624] * 4
    if q == 3:
        a, b, c, d = 0.77849529482132992, 0.0738349017262233984, 0.406244343884051008, 0.0937556561159491072
        pts = [[a, b, b], [b, a, b], [b, b, a], [b, b, b], [c, c, d], [c, d, c], [c, d, d], [d, c, c], [d, c, d],
               [d, d, c]]
        return pts, [0.00793885580720148224] * 4 + [0.022485207239643504] * 6
    if q == 4:
        a, b = 0.902942215818267904, 0.0323525947272438976
        c, d, e = 0.262682583887778976, 0.616596533061936896, 0.0603604415251420928
        f, g = 0.309769304272861952, 0.0706920871814129024
        pts = [[a, b, b], [b, a, b], [b, b, a], [b, b, b], [c, d, e], [d, c, e], [c, e, d], [d, e, c], [c, e, e],
               [d, e, e], [e, c, d], [e, d, c], [e, c, e], [e, d, e], [e, e, c], [e, e, d], [f, f, f], [f, f, g],
               [f, g, f], [g, f, f]]
        return pts, [0.00117784579907825008] * 4 + [0.0078331114953146176] * 12 + [0.0169894863816446688] * 4
    raise ValueError(q)


def _tri_rule(q):
    if q == 1:
        return [[0.333333333333333, 0.333333333333333]], [0.5]
    if q == 2:
        a, b = 0.166666666666667, 0.666666666666667
        return [[a, a], [a, b], [b, a]], [0.166666666666667] * 3
    if q == 3:
        a, b, c, d = 0.09157621350978, 0.81684757298044, 0.108103018168071, 0.445948490915964
        return [[a, a], [a, b], [b, a], [c, d], [d, c], [d, d]], [0.0549758718276665] * 3 + [0.111690794839] * 3
    if q == 4:
        a, b = 0.055564052669793, 0.888871894660413
        c, d, e = 0.070255540518384, 0.634210747745723, 0.295533711735893
        t = 0.333333333333333
        pts = [[a, a], [a, b], [b, a], [c, d], [d, c], [d, e], [c, e], [e, c], [e, d], [t, t]]
        return pts, [0.0209777564983245] * 3 + [0.0560492060354435] * 6 + [0.100771494292365]
    raise ValueError(q)


def is_simplex(geom):
    return geom in (TRIANGLE, TETRAHEDRON)


def num_quadrature_points(geom, q):
    """F/geometry.hpp:25-43"""
    return {SEGMENT: q, TRIANGLE: q * (q + 1) // 2, SQUARE: q * q, TETRAHEDRON: q * (q + 1) * (q + 2) // 6,
            CUBE: q**3}[geom]


def simplex_quadrature(geom, q):
    pts, w = _tet_rule(q) if geom == TETRAHEDRON else _tri_rule(q)
    return np.array(pts), np.array(w)


def simplex_lattice(geom, p):
    """integer lattice coordinates of the nodes in the reference's (lexicographic, x fastest) numbering"""
    if geom == TRIANGLE:
        return [(i, j) for j in range(p + 1) for i in range(p + 1 - j)]
    return [(i, j, k) for k in range(p + 1) for j in range(p + 1 - k) for i in range(p + 1 - k - j)]


def simplex_nodes(geom, p):
    """node positions: Gauss-Lobatto points along edges, centroids inside faces (p = 3), as the nodal polynomials of
    triangle_H1.inl / tetrahedron_H1.inl imply (each is 1 at its own node and 0 at the others)"""
    gl = {1: [0.0, 1.0], 2: [0.0, 0.5, 1.0], 3: [0.0, 0.5 - 0.1 * _SQRT5, 0.5 + 0.1 * _SQRT5, 1.0]}[p]
    out = []
    for c in simplex_lattice(geom, p):
        last = p - sum(c)
        bary = list(c) + [last]
        nz = sum(1 for b in bary if b > 0)
        if nz == 3:
            x = [1.0 / 3.0 if b > 0 else 0.0 for b in c]
        else:
            x = [gl[b] for b in c]
            if nz == 2 and last == 0:  # edge opposite to the origin: coordinates sum to one
                i0, i1 = [d for d, b in enumerate(c) if b > 0]
                x[i1] = 1.0 - x[i0]
        out.append(x)
    return np.array(out)


def simplex_shape_tables(geom, p, pts):
    """phi [npts][ndof], dphi [npts][ndof][dim] of the nodal basis of P_p.  p = 1, 2: the closed forms of
    triangle_H1.inl:50-80 / tetrahedron_H1.inl:47-84; p = 3: the unique Lagrange basis through simplex_nodes (the
    reference spells the same polynomials out), from the monomial Vandermonde system."""
    pts = np.asarray(pts, dtype=np.float64)
    dim = pts.shape[1]
    x = [pts[:, d] for d in range(dim)]
    one, zero = np.ones(len(pts)), np.zeros(len(pts))
    if p == 0:  # L2<0> / QOI: one shape function, identically 1
        return np.ones((len(pts), 1)), np.zeros((len(pts), 1, dim))
    if p == 1:
        lam0 = 1.0 - sum(x)
        phi = [lam0] + x
        dphi = [[-one] * dim] + [[one if d == e else zero for e in range(dim)] for d in range(dim)]
        return np.stack(phi, 1), np.stack([np.stack(g, 1) for g in dphi], 1)
    if p == 2:
        lam = [1.0 - sum(x)] + x  # barycentric coordinates; node (i,j,k) <-> pair of vertices
        dlam = [[-1.0] * dim] + [[1.0 if d == e else 0.0 for e in range(dim)] for d in range(dim)]
        phi, dphi = [], []
        for c in simplex_lattice(geom, 2):
            bary = [2 - sum(c)] + list(c)
            idx = [v for v, b in enumerate(bary) for _ in range(b)]  # the two (possibly equal) vertices
            a, b = idx
            if a == b:  # vertex function lam (2 lam - 1)
                phi.append(lam[a] * (2.0 * lam[a] - 1.0))
                dphi.append(np.stack([(4.0 * lam[a] - 1.0) * dlam[a][e] for e in range(dim)], 1))
            else:  # edge function 4 lam_a lam_b
                phi.append(4.0 * lam[a] * lam[b])
                dphi.append(np.stack([4.0 * (lam[a] * dlam[b][e] + lam[b] * dlam[a][e]) for e in range(dim)], 1))
        return np.stack(phi, 1), np.stack(dphi, 1)
    nodes = simplex_nodes(geom, p)
    expo = simplex_lattice(geom, p)  # monomial exponents: same index set

    def monomials(P, deriv=None):
        cols = []
        for ex in expo:
            v = np.ones(len(P))
            for d in range(dim):
                e = ex[d]
                if d == deriv:
                    if e == 0:
                        v = np.zeros(len(P))
                        break
                    v = v * e
                    e -= 1
                v = v * P[:, d]**e
            cols.append(v)
        return np.stack(cols, 1)

    C = np.linalg.inv(monomials(nodes))  # phi_k = sum_m C[m, k] mono_m
    phi = monomials(pts) @ C
    dphi = np.stack([monomials(pts, d) @ C for d in range(dim)], 2)
    return phi, dphi


# ----------------------------------------------------------------------------------------------------------------------
# ElementRestriction (F/element_restriction.cpp)
# ----------------------------------------------------------------------------------------------------------------------
def lexicographic_permutation(geom, p):
    """native (mfem H1: vertices, edges, faces, interior; mfem/fem/fe/fe_h1.cpp) -> lexicographic, i.e. what
    H1_*Element(p).GetLexicographicOrdering() returns (F/element_restriction.cpp:7-65).  lex[k] = native index of the
    node at lexicographic position k.  Built from node coordinates, independently of the product's tables."""
    n = p + 1
    interior = list(range(1, p))
    if geom == SEGMENT:
        native = [(0,), (p,)] + [(i,) for i in interior]
    elif geom == SQUARE:
        native = [(0, 0), (p, 0), (p, p), (0, p)]
        native += [(i, 0) for i in interior] + [(p, i) for i in interior]
        native += [(p - i, p) for i in interior] + [(0, p - i) for i in interior]
        native += [(i, j) for j in interior for i in interior]
    elif geom == CUBE:
        native = [(0, 0, 0), (p, 0, 0), (p, p, 0), (0, p, 0), (0, 0, p), (p, 0, p), (p, p, p), (0, p, p)]
        for z in (0, p):  # edges 0-3 (z = 0) and 4-7 (z = p)
            native += [(i, 0, z) for i in interior] + [(p, i, z) for i in interior]
            native += [(i, p, z) for i in interior] + [(0, i, z) for i in interior]
        for (x, y) in ((0, 0), (p, 0), (p, p), (0, p)):  # vertical edges 8-11
            native += [(x, y, i) for i in interior]
        native += [(i, p - j, 0) for j in interior for i in interior]  # face 0: z = 0
        native += [(i, 0, j) for j in interior for i in interior]      # face 1: y = 0
        native += [(p, i, j) for j in interior for i in interior]      # face 2: x = p
        native += [(p - i, p, j) for j in interior for i in interior]  # face 3: y = p
        native += [(0, p - i, j) for j in interior for i in interior]  # face 4: x = 0
        native += [(i, j, p) for j in interior for i in interior]      # face 5: z = p
        native += [(i, j, k) for k in interior for j in interior for i in interior]
    elif geom in (TRIANGLE, TETRAHEDRON):
        # mfem H1_TriangleElement / H1_TetrahedronElement native order: vertices, edges, faces, interior
        if geom == TRIANGLE:
            native = [(0, 0), (p, 0), (0, p)]
            native += [(i, 0) for i in interior] + [(p - i, i) for i in interior] + [(0, p - i) for i in interior]
            native += [(i, j) for j in interior for i in range(1, p - j)]
        else:
            native = [(0, 0, 0), (p, 0, 0), (0, p, 0), (0, 0, p)]
            native += [(i, 0, 0) for i in interior] + [(0, i, 0) for i in interior] + [(0, 0, i) for i in interior]
            native += [(p - i, i, 0) for i in interior] + [(p - i, 0, i) for i in interior]
            native += [(0, p - i, i) for i in interior]
            face_ij = [(i, j) for j in interior for i in range(1, p - j)]
            native += [(p - i - j, i, j) for (i, j) in face_ij] + [(0, j, i) for (i, j) in face_ij]
            native += [(i, 0, j) for (i, j) in face_ij] + [(j, i, 0) for (i, j) in face_ij]
            native += [(i, j, k) for k in interior for j in range(1, p - k) for i in range(1, p - k - j)]
        lattice = simplex_lattice(geom, p)
        lex = np.empty(len(native), dtype=np.int64)
        for nat, c in enumerate(native):
            lex[lattice.index(c)] = nat
        return lex
    else:
        raise ValueError("unsupported geometry")
    lex = np.empty(len(native), dtype=np.int64)
    for nat, c in enumerate(native):
        k = 0
        for d in reversed(range(len(c))):
            k = k * n + c[d]
        lex[k] = nat
    return lex


def dof_info(geom, p, native_dofs):
    """GetElementRestriction for H1 (F/element_restriction.cpp:216-277): dof_info(e,k) = dofs[lex_perm[k]], packed as a
    DoF word with zero sign/orientation bits (F/element_restriction.hpp:31-80)."""
    lex = lexicographic_permutation(geom, p)
    return np.asarray(native_dofs)[:, lex].astype(np.uint64) & np.uint64(0x0000FFFFFFFFFFFF)


def vdof(node, comp, components, num_nodes, byvdim=True):
    """ElementRestriction::GetVDof (F/element_restriction.cpp:425-432)"""
    return node * components + comp if byvdim else comp * num_nodes + node


def gather(table, L, components, byvdim=True):
    """ElementRestriction::Gather (F/element_restriction.cpp:443-454) -> E[e][c][k]"""
    table = np.asarray(table).astype(np.int64)
    num_nodes = len(L) // components
    idx = np.stack([vdof(table, c, components, num_nodes, byvdim) for c in range(components)], axis=1)
    return np.asarray(L)[idx]


def scatter_add(table, E, L, components, byvdim=True):
    """ElementRestriction::ScatterAdd (F/element_restriction.cpp:456-467); np.add.at applies the additions one by one
    in index order, i.e. in the reference's element-major order, so the sums are bit-identical."""
    table = np.asarray(table).astype(np.int64)
    num_nodes = len(L) // components
    idx = np.stack([vdof(table, c, components, num_nodes, byvdim) for c in range(components)], axis=1)
    np.add.at(L, idx.ravel(), np.asarray(E).ravel())
    return L


# ----------------------------------------------------------------------------------------------------------------------
# tensor-product H1 element: interpolate / integrate (F/detail/hexahedron_H1.inl:169-280, quadrilateral_H1.inl)
# ----------------------------------------------------------------------------------------------------------------------
def interpolate(u_E, p, q, dim, geom=None):
    """u_E [ne][c][n^dim] (lexicographic, x fastest) -> value [ne][c][nq], parent gradient [ne][c][dim][nq] with the
    quadrature index Q = (qz*q + qy)*q + qx.  Simplex geometries: dense sums over the nodal basis
    (triangle_H1.inl:269-292, tetrahedron_H1.inl:369-390)."""
    if geom is not None and is_simplex(geom):
        xi, _ = simplex_quadrature(geom, q)
        phi, dphi = simplex_shape_tables(geom, p, xi)
        return np.einsum("qk,eck->ecq", phi, u_E), np.einsum("qkj,eck->ecjq", dphi, u_E)
    B, G, _ = basis_tables(p, q)
    ne, c = u_E.shape[:2]
    n = p + 1
    if dim == 3:
        X = u_E.reshape(ne, c, n, n, n)  # [z][y][x]
        val = np.einsum("wk,vj,ui,eckji->ecwvu", B, B, B, X)
        gx = np.einsum("wk,vj,ui,eckji->ecwvu", B, B, G, X)
        gy = np.einsum("wk,vj,ui,eckji->ecwvu", B, G, B, X)
        gz = np.einsum("wk,vj,ui,eckji->ecwvu", G, B, B, X)
        grad = np.stack([gx, gy, gz], axis=2)
    elif dim == 2:
        X = u_E.reshape(ne, c, n, n)
        val = np.einsum("vj,ui,ecji->ecvu", B, B, X)
        gx = np.einsum("vj,ui,ecji->ecvu", B, G, X)
        gy = np.einsum("vj,ui,ecji->ecvu", G, B, X)
        grad = np.stack([gx, gy], axis=2)
    else:  # segment (F/detail/segment_H1.inl:86-100)
        val = np.einsum("ui,eci->ecu", B, u_E)
        grad = np.einsum("ui,eci->ecu", G, u_E)[:, :, None, :]
    nq = q**dim
    return val.reshape(ne, c, nq), grad.reshape(ne, c, dim, nq)


def integrate(source, flux, p, q, dim, geom=None):
    """source [ne][c][nq] or None, flux [ne][c][dim][nq] or None (parent space) -> r_E [ne][c][n^dim];
    weights folded in as in hexahedron_H1.inl:244-246.  Simplex geometries: tetrahedron_H1.inl:392-429."""
    if geom is not None and is_simplex(geom):
        xi, w = simplex_quadrature(geom, q)
        phi, dphi = simplex_shape_tables(geom, p, xi)
        ref = source if source is not None else flux
        out = np.zeros(ref.shape[:2] + (phi.shape[1],))
        if source is not None:
            out += np.einsum("ecq,qk,q->eck", source, phi, w)
        if flux is not None:
            out += np.einsum("ecjq,qkj,q->eck", flux, dphi, w)
        return out
    B, G, w = basis_tables(p, q)
    Bw, Gw = B * w[:, None], G * w[:, None]
    n = p + 1
    ref = source if source is not None else flux
    ne, c = ref.shape[:2]
    out = np.zeros((ne, c) + (n,) * dim)
    if dim == 3:
        if source is not None:
            out += np.einsum("wk,vj,ui,ecwvu->eckji", Bw, Bw, Bw, source.reshape(ne, c, q, q, q))
        if flux is not None:
            f = flux.reshape(ne, c, 3, q, q, q)
            out += np.einsum("wk,vj,ui,ecwvu->eckji", Bw, Bw, Gw, f[:, :, 0])
            out += np.einsum("wk,vj,ui,ecwvu->eckji", Bw, Gw, Bw, f[:, :, 1])
            out += np.einsum("wk,vj,ui,ecwvu->eckji", Gw, Bw, Bw, f[:, :, 2])
    elif dim == 2:
        if source is not None:
            out += np.einsum("vj,ui,ecvu->ecji", Bw, Bw, source.reshape(ne, c, q, q))
        if flux is not None:
            f = flux.reshape(ne, c, 2, q, q)
            out += np.einsum("vj,ui,ecvu->ecji", Bw, Gw, f[:, :, 0])
            out += np.einsum("vj,ui,ecvu->ecji", Gw, Bw, f[:, :, 1])
    else:  # segment (F/detail/segment_H1.inl:102-118)
        if source is not None:
            out += np.einsum("ui,ecu->eci", Bw, source)
        if flux is not None:
            out += np.einsum("ui,ecu->eci", Gw, flux[:, :, 0])
    return out.reshape(ne, c, n**dim)


def geometric_factors(nodes_E, mesh_order, q, dim, geom=None):
    """compute_geometric_factors (F/geometric_factors.cpp:16-60): X_q [ne][dim][nq], J_q [ne][j][i][nq] = dX_i/dxi_j"""
    val, grad = interpolate(nodes_E, mesh_order, q, dim, geom)
    return val, np.transpose(grad, (0, 2, 1, 3))


# ----------------------------------------------------------------------------------------------------------------------
# forward-mode AD on arrays
# ----------------------------------------------------------------------------------------------------------------------
class Dual:
    """value v [...], partials g [..., N]"""
    __array_priority__ = 100

    def __init__(self, v, g):
        self.v, self.g = v, g

    @staticmethod
    def lift(x, like):
        return x if isinstance(x, Dual) else Dual(np.broadcast_to(x, like.v.shape), np.zeros_like(like.g))

    def __add__(self, o):
        return Dual(self.v + o.v, self.g + o.g) if isinstance(o, Dual) else Dual(self.v + o, self.g)

    __radd__ = __add__

    def __neg__(self):
        return Dual(-self.v, -self.g)

    def __sub__(self, o):
        return Dual(self.v - o.v, self.g - o.g) if isinstance(o, Dual) else Dual(self.v - o, self.g)

    def __rsub__(self, o):
        return Dual(o - self.v, -self.g)

    def __mul__(self, o):
        if isinstance(o, Dual):
            return Dual(self.v * o.v, self.v[..., None] * o.g + self.g * o.v[..., None])
        o = np.asarray(o)
        return Dual(self.v * o, self.g * (o[..., None] if o.ndim else o))

    __rmul__ = __mul__

    def __truediv__(self, o):
        if isinstance(o, Dual):
            q = self.v / o.v
            return Dual(q, (self.g - q[..., None] * o.g) / o.v[..., None])
        o = np.asarray(o)
        return Dual(self.v / o, self.g / (o[..., None] if o.ndim else o))

    def __rtruediv__(self, o):
        q = o / self.v
        return Dual(q, -(q / self.v)[..., None] * self.g)


def value_of(x):
    return x.v if isinstance(x, Dual) else x


def _chain(x, f, df):
    return Dual(f, df[..., None] * x.g) if isinstance(x, Dual) else f


def d_log1p(x):
    v = value_of(x)
    return _chain(x, np.log1p(v), 1.0 / (1.0 + v))


def d_sqrt(x):
    v = value_of(x)
    s = np.sqrt(v)
    return _chain(x, s, 0.5 / s)


# 3x3 (or dxd) tensors are lists of lists of scalars (numpy arrays or Duals)
def eye(d):
    return [[1.0 if i == j else 0.0 for j in range(d)] for i in range(d)]


def madd(A, B):
    return [[A[i][j] + B[i][j] for j in range(len(A[0]))] for i in range(len(A))]


def msub(A, B):
    return [[A[i][j] - B[i][j] for j in range(len(A[0]))] for i in range(len(A))]


def mscale(s, A):
    return [[s * A[i][j] for j in range(len(A[0]))] for i in range(len(A))]


def mdot(A, B):
    return [[sum(A[i][k] * B[k][j] for k in range(len(B))) for j in range(len(B[0]))] for i in range(len(A))]


def mT(A):
    return [[A[j][i] for j in range(len(A))] for i in range(len(A[0]))]


def mtr(A):
    return sum(A[i][i] for i in range(len(A)))


def mdet(A):
    if len(A) == 2:
        return A[0][0] * A[1][1] - A[0][1] * A[1][0]
    return (A[0][0] * A[1][1] * A[2][2] + A[0][1] * A[1][2] * A[2][0] + A[0][2] * A[1][0] * A[2][1] -
            A[0][0] * A[1][2] * A[2][1] - A[0][1] * A[1][0] * A[2][2] - A[0][2] * A[1][1] * A[2][0])


def minv(A):
    """F/tensor.hpp:1586-1632 (closed forms; for duals the reference solves A X = I, which is the same inverse)"""
    d = mdet(A)
    if len(A) == 2:
        return [[A[1][1] / d, -1.0 * A[0][1] / d], [-1.0 * A[1][0] / d, A[0][0] / d]]
    c = [[A[1][1] * A[2][2] - A[1][2] * A[2][1], A[0][2] * A[2][1] - A[0][1] * A[2][2], A[0][1] * A[1][2] - A[0][2] * A[1][1]],
         [A[1][2] * A[2][0] - A[1][0] * A[2][2], A[0][0] * A[2][2] - A[0][2] * A[2][0], A[0][2] * A[1][0] - A[0][0] * A[1][2]],
         [A[1][0] * A[2][1] - A[1][1] * A[2][0], A[0][1] * A[2][0] - A[0][0] * A[2][1], A[0][0] * A[1][1] - A[0][1] * A[1][0]]]
    return [[c[i][j] / d for j in range(3)] for i in range(3)]


def detApIm1(A):
    """F/tensor.hpp:1266-1296"""
    if len(A) == 2:
        return A[0][0] - A[0][1] * A[1][0] + A[1][1] + A[0][0] * A[1][1]
    return (A[0][0] + A[1][1] + A[2][2] - A[0][1] * A[1][0] * (1 + A[2][2]) + A[0][0] * A[1][1] * (1 + A[2][2]) -
            A[0][2] * A[2][0] * (1 + A[1][1]) - A[1][2] * A[2][1] * (1 + A[0][0]) + A[0][0] * A[2][2] + A[1][1] * A[2][2] +
            A[0][1] * A[1][2] * A[2][0] + A[0][2] * A[1][0] * A[2][1])


def msym(A):
    return [[0.5 * (A[i][j] + A[j][i]) for j in range(len(A))] for i in range(len(A))]


def mdev(A):
    t = mtr(A) / float(len(A))
    return [[A[i][j] - (t if i == j else 0.0) for j in range(len(A))] for i in range(len(A))]


def mnorm(A):
    return d_sqrt(sum(A[i][j] * A[i][j] for i in range(len(A)) for j in range(len(A[0]))))


# ----------------------------------------------------------------------------------------------------------------------
# material models and the functors the physics modules wrap them in
# ----------------------------------------------------------------------------------------------------------------------
class NeoHookean:
    """src/serac/physics/materials/solid_material.hpp:98-127"""

    def __init__(self, density, K, G):
        self.density, self.K, self.G = density, K, G

    def __call__(self, state, du_dX):
        d = len(du_dX)
        lam = self.K - (2.0 / 3.0) * self.G
        B_minus_I = madd(madd(mdot(du_dX, mT(du_dX)), mT(du_dX)), du_dX)
        J_minus_1 = detApIm1(du_dX)
        J = J_minus_1 + 1
        ll = lam * d_log1p(J_minus_1)
        return [[(ll * (1.0 if i == j else 0.0) + self.G * B_minus_I[i][j]) / J for j in range(d)] for i in range(d)]


class LinearIsotropic:
    """src/serac/physics/materials/solid_material.hpp:24-50"""

    def __init__(self, density, K, G):
        self.density, self.K, self.G = density, K, G

    def __call__(self, state, du_dX):
        d = len(du_dX)
        lam = self.K - (2.0 / 3.0) * self.G
        eps = mscale(0.5, madd(mT(du_dX), du_dX))
        t = mtr(eps)
        return [[lam * t * (1.0 if i == j else 0.0) + 2.0 * self.G * eps[i][j] for j in range(d)] for i in range(d)]


class StVenantKirchhoff:
    """src/serac/physics/materials/solid_material.hpp:61-92"""

    def __init__(self, density, K, G):
        self.density, self.K, self.G = density, K, G

    def __call__(self, state, grad_u):
        d = len(grad_u)
        F = madd(grad_u, eye(d))
        E = mscale(0.5, madd(madd(grad_u, mT(grad_u)), mdot(mT(grad_u), grad_u)))
        trE = mtr(E)
        devE = mdev(E)
        S = [[2.0 * self.G * devE[i][j] + (self.K * trE if i == j else 0.0) for j in range(d)] for i in range(d)]
        P = mdot(F, S)
        return mscale(1.0 / mdet(F), mdot(P, mT(F)))


class J2SmallStrainLinear:
    """src/serac/physics/materials/solid_material.hpp:201-259 with LinearHardening (:132-149).  The return map's scalar
    residual is linear in delta_eqps for linear hardening, so the safeguarded Newton solve of
    tuple_tensor_dual_functions.hpp:570-677 converges in one step to the closed form used here; the dual part follows
    the implicit-function rule at :664-676.  state = dict(plastic_strain [..][3][3] list of arrays, eqps array)."""
    tol = 1e-10

    def __init__(self, E, nu, sigma_y, Hi, Hk, density, hardening=None):
        self.E, self.nu, self.sigma_y, self.Hi, self.Hk, self.density = E, nu, sigma_y, Hi, Hk, density
        # hardening law: (value, derivative) of the flow stress as functions of the accumulated plastic strain
        self.hardening = hardening or ((lambda e: sigma_y + Hi * e), (lambda e: Hi + 0.0 * e))

    def __call__(self, state, du_dX):
        K = self.E / (3.0 * (1.0 - 2.0 * self.nu))
        G = 0.5 * self.E / (1.0 + self.nu)
        ep, eqps_old = state["plastic_strain"], state["eqps"]
        el = msub(msym(du_dX), ep)
        pr = K * mtr(el)
        s = mscale(2.0 * G, mdev(el))
        sigma_b = mscale(2.0 / 3.0 * self.Hk, ep)
        eta = msub(s, sigma_b)
        q = np.sqrt(1.5) * mnorm(eta)
        qv = value_of(q)
        h, dh = self.hardening
        resid0 = qv - h(eqps_old)
        yielding = resid0 > self.tol * self.sigma_y
        # root of  q - (3G + Hk) d - h(eqps_old + d) = 0  (solid_material.hpp:229-247; Newton on the values -- one step for
        # linear hardening), then d(delta)/dq = 1 / (3G + Hk + h'(eqps_old + d)) by the implicit-function rule
        # (tuple_tensor_dual_functions.hpp:664-676)
        dvn = np.zeros_like(qv)
        for _ in range(60):
            f = qv - (3.0 * G + self.Hk) * dvn - h(eqps_old + dvn)
            dvn = np.where(yielding, dvn + f / (3.0 * G + self.Hk + dh(eqps_old + dvn)), 0.0)
        slope = 3.0 * G + self.Hk + dh(eqps_old + dvn)
        if isinstance(q, Dual):
            delta = Dual(dvn, q.g / slope[..., None])
        else:
            delta = dvn
        mask = yielding.astype(float)
        qsafe = q if not isinstance(q, Dual) else Dual(np.where(yielding, q.v, 1.0), q.g)
        if not isinstance(q, Dual):
            qsafe = np.where(yielding, q, 1.0)
        Np = [[1.5 * eta[i][j] / qsafe for j in range(3)] for i in range(3)]
        s = [[s[i][j] - mask * (2.0 * G * delta * Np[i][j]) for j in range(3)] for i in range(3)]
        dv = value_of(delta) * mask
        new_state = dict(eqps=eqps_old + dv,
                         plastic_strain=[[ep[i][j] + dv * value_of(Np[i][j]) for j in range(3)] for i in range(3)])
        state["_updated"] = new_state
        return [[s[i][j] + (pr if i == j else 0.0) for j in range(3)] for i in range(3)]


def _stack_values(A):
    return np.stack([np.stack([np.broadcast_arrays(value_of(A[i][j]), value_of(A[0][0]))[0] for j in range(3)], -1)
                     for i in range(3)], -2)


def symmetric_mat3_function(A, f, g):
    """f(A) = Q diag(f(lambda)) Q^T for a symmetric 3x3 tensor given as a list of lists of arrays / Duals; with Dual
    entries the derivative is sum_ab g(l_a, l_b) Q_ka Q_lb Q_ia Q_jb dA_kl  (F/tuple_tensor_dual_functions.hpp:963-1000).
    The eigen-decomposition is numpy's (the reference uses the closed form of :802-912; f(A) and its derivative do not
    depend on which orthonormal eigenbasis is picked)."""
    Av = _stack_values(A)
    lam, Q = np.linalg.eigh(Av)
    fA = np.einsum("...ia,...a,...ja->...ij", Q, f(lam), Q)
    duals = [A[i][j] for i in range(3) for j in range(3) if isinstance(A[i][j], Dual)]
    if not duals:
        return [[fA[..., i, j] for j in range(3)] for i in range(3)]
    N = duals[0].g.shape[-1]
    zero = np.zeros(Av.shape[:-2] + (N,))
    Ag = np.stack([np.stack([A[i][j].g if isinstance(A[i][j], Dual) else zero for j in range(3)], -2)
                   for i in range(3)], -3)  # [...][k][l][N]
    G = g(lam[..., :, None], lam[..., None, :])
    B = np.einsum("...ka,...kln,...lb->...abn", Q, Ag, Q) * G[..., None]
    out_g = np.einsum("...ia,...abn,...jb->...ijn", Q, B, Q)
    return [[Dual(fA[..., i, j], out_g[..., i, j, :]) for j in range(3)] for i in range(3)]


def log_symm(A):
    """F/tuple_tensor_dual_functions.hpp:1007-1020"""
    def g(l1, l2):
        same = l1 == l2
        y = np.where(same, 2.0, l1 / l2)
        return np.where(same, 1.0 / l1, (np.log(y) / (y - 1.0)) / l2)
    return symmetric_mat3_function(A, np.log, g)


def exp_symm(A):
    """F/tuple_tensor_dual_functions.hpp:1028-1041"""
    def g(l1, l2):
        same = l1 == l2
        arg = np.where(same, 1.0, l1 - l2)
        return np.where(same, np.exp(l1), np.exp(l2) * np.expm1(arg) / arg)
    return symmetric_mat3_function(A, np.exp, g)


class J2FiniteLinear:
    """J2<LinearHardening> (src/serac/physics/materials/solid_material.hpp:261-330): finite-deformation J2 in logarithmic
    strain; state = {"Fpinv": 3x3, "eqps"}.  Linear hardening makes the return-map equation linear, so its root is
    written in closed form (the reference's scalar solver converges to the same root)."""
    tol = 1e-10

    def __init__(self, E, nu, sigma_y, Hi, density):
        self.E, self.nu, self.sigma_y, self.Hi, self.density = E, nu, sigma_y, Hi, density

    def __call__(self, state, du_dX):
        K = self.E / (3.0 * (1.0 - 2.0 * self.nu))
        G = 0.5 * self.E / (1.0 + self.nu)
        Fpinv, eqps_old = state["Fpinv"], state["eqps"]
        F = madd(du_dX, eye(3))
        Fe = mdot(F, Fpinv)
        Ee = mscale(0.5, log_symm(mdot(mT(Fe), Fe)))
        pr = K * mtr(Ee)
        s = mscale(2.0 * G, mdev(Ee))
        q = np.sqrt(1.5) * mnorm(s)
        qv = value_of(q)
        yielding = (qv - (self.sigma_y + self.Hi * eqps_old)) > self.tol * self.sigma_y
        delta = (q - (self.sigma_y + self.Hi * eqps_old)) / (3.0 * G + self.Hi)
        mask = yielding.astype(float)
        if isinstance(q, Dual):
            qsafe = Dual(np.where(yielding, q.v, 1.0), q.g)
        else:
            qsafe = np.where(yielding, q, 1.0)
        Np = [[1.5 * s[i][j] / qsafe for j in range(3)] for i in range(3)]
        dm = delta * mask  # zero (value and partials) where the step is elastic
        s = [[s[i][j] - 2.0 * G * dm * Np[i][j] for j in range(3)] for i in range(3)]
        A = exp_symm([[-1.0 * (dm * Np[i][j]) for j in range(3)] for i in range(3)])  # identity where elastic
        Fe = mdot(Fe, A)
        Av = [[value_of(A[i][j]) for j in range(3)] for i in range(3)]
        state["_updated"] = dict(eqps=eqps_old + value_of(dm), Fpinv=mdot(Fpinv, Av))
        M = [[s[i][j] + (pr if i == j else 0.0) for j in range(3)] for i in range(3)]
        FeT = mT(Fe)
        sig = mdot(mdot(minv(FeT), M), FeT)
        return mscale(1.0 / mdet(F), sig)


class LinearIsotropicConductor:
    """src/serac/physics/materials/thermal_material.hpp:20-65"""

    def __init__(self, density, specific_heat, conductivity):
        self.rho, self.cp, self.k = density, specific_heat, conductivity

    def __call__(self, x, u, du_dX):
        return self.rho * self.cp, [-1.0 * self.k * g for g in du_dX]


class GreenSaintVenantThermoelastic:
    """src/serac/physics/materials/green_saint_venant_thermoelastic.hpp:26-92: (state, grad_u, theta, grad_theta) ->
    (Cauchy stress, C_v, internal heat source, heat flux); state = {"strain_trace"}"""

    def __init__(self, density, E, nu, C_v, alpha, theta_ref, k):
        self.density, self.E, self.nu, self.C_v, self.alpha, self.theta_ref, self.k = density, E, nu, C_v, alpha, theta_ref, k

    def __call__(self, state, grad_u, theta, grad_theta):
        K = self.E / (3.0 * (1.0 - 2.0 * self.nu))
        G = 0.5 * self.E / (1.0 + self.nu)
        F = madd(grad_u, eye(3))
        Eg = mscale(0.5, madd(madd(grad_u, mT(grad_u)), mdot(mT(grad_u), grad_u)))
        trEg = mtr(Eg)
        vol = K * (trEg - 3.0 * self.alpha * (theta - self.theta_ref))
        devE = mdev(Eg)
        S = [[2.0 * G * devE[i][j] + (vol if i == j else 0.0) for j in range(3)] for i in range(3)]
        P = mdot(F, S)
        sigma = mscale(1.0 / mdet(F), mdot(P, mT(F)))
        s0 = -3.0 * K * self.alpha * theta * (trEg - state["strain_trace"])
        q0 = [-1.0 * self.k * g for g in grad_theta]
        state["_updated"] = dict(strain_trace=value_of(trEg))
        return sigma, self.C_v, s0, q0


def thermoelastic_solid_qf(material, geom_nonlin):
    """MaterialStressFunctor<MechanicalMaterialInterface<Material>> (solid_mechanics.hpp:834-852 with the parameter pack,
    thermomechanics.hpp:318-352): arguments (displacement, acceleration, temperature)"""

    def qf(t, x, state, displacement, acceleration, temperature):
        du_dX = displacement[1]
        stress = material(state, du_dX, temperature[0][0], temperature[1][0])[0]
        dx_dX = madd(mscale(0.0, du_dX), eye(3))
        if geom_nonlin:
            dx_dX = madd(dx_dX, du_dX)
        flux = mscale(mdet(dx_dX), mdot(stress, mT(minv(dx_dX))))
        return [material.density * a for a in acceleration[0]], flux

    return qf


def thermoelastic_thermal_qf(material):
    """ThermalMaterialIntegrand<ThermalMaterialInterface<Material>> (heat_transfer.hpp:392-402, thermomechanics.hpp:272-310):
    arguments (temperature, dT/dt, displacement); the material's state is a throw-away local, as in the reference"""

    def qf(t, x, state, temperature, dtemp_dt, displacement):
        zero = 0.0 * value_of(temperature[0][0])
        local = dict(strain_trace=zero)
        _, cap, _, q0 = material(local, displacement[1], temperature[0][0], temperature[1][0])
        return [cap * dtemp_dt[0][0]], [[-1.0 * h for h in q0]]

    return qf


def material_stress_qf(material, geom_nonlin):
    """SolidMechanics::MaterialStressFunctor (src/serac/physics/solid_mechanics.hpp:808-853):
    (state, displacement=(u, du_dX), acceleration=(a, da_dX)) -> (source [c], flux [c][dim])"""

    def qf(t, x, state, displacement, acceleration):
        du_dX = displacement[1]
        d = len(du_dX)
        stress = material(state, du_dX)
        dx_dX = madd(mscale(0.0, du_dX), eye(d))
        if geom_nonlin:
            dx_dX = madd(dx_dX, du_dX)
        flux = mscale(mdet(dx_dX), mdot(stress, mT(minv(dx_dX))))
        return [material.density * a for a in acceleration[0]], flux

    return qf


def thermal_qf(material):
    """HeatTransfer::ThermalMaterialIntegrand (src/serac/physics/heat_transfer.hpp:381-405); scalar field: c = 1"""

    def qf(t, x, state, temperature, dtemp_dt):
        u, du_dX = temperature[0][0], temperature[1][0]
        cap, heat_flux = material(x, u, du_dX)
        return [cap * dtemp_dt[0][0]], [[-1.0 * h for h in heat_flux]]

    return qf


def shape_aware(qf):
    """ShapeAwareIntegrandWrapper[WithState] (F/shape_aware_functional.hpp:364-438) with detail::ShapeCorrection (:37-141)
    and apply_shape_aware_qf_helper (:199-217): argument 0 is the shape displacement p, x' = X + p."""

    def wrapped(t, x, state, shape, *args):
        X, dX_dxi = x
        d = len(dX_dxi)
        J = madd(eye(d), shape[1])
        detJ, invJ = mdet(J), minv(J)
        invJT = mT(invJ)
        x_mod = ([X[i] + shape[0][i] for i in range(d)], madd(dX_dxi, mdot(shape[1], dX_dxi)))
        mod_args = [(a[0], mdot(a[1], invJ)) for a in args]  # modify_trial_argument: du/dx' = du/dX . inv(J)
        source, flux = qf(t, x_mod, state, *mod_args)
        return [s * detJ for s in source], mscale(detJ, mdot(flux, invJT))  # modify_shape_aware_qf_return

    return wrapped


# ----------------------------------------------------------------------------------------------------------------------
# the three kernel families of an Integral on tensor-product elements (F/domain_integral_kernels.hpp)
# ----------------------------------------------------------------------------------------------------------------------
def _tensor_lists(val, grad):
    """[ne][c][nq], [ne][c][dim][nq] -> per-quadrature-point lists with arrays of shape [ne][nq]"""
    c, dim = grad.shape[1], grad.shape[2]
    return [val[:, a] for a in range(c)], [[grad[:, a, j] for j in range(dim)] for a in range(c)]


def evaluation_kernel(qf, trials_E, orders, X_q, J_q, test_order, q, dim, wrt=-1, state=None, geom=None):
    """evaluation_kernel_impl (F/domain_integral_kernels.hpp:152-220).  trials_E: list of [ne][c][ndof] E-vectors.
    Returns r_E [ne][ct][ndof] and, when wrt >= 0, the derivative blocks after physical_to_parent:
      d00 [ne][nq][ct][ck], d01 [ne][nq][ct][ck][dim], d10 [ne][nq][ct][dim][ck], d11 [ne][nq][ct][dim][ck][dim]."""
    ne, nq = X_q.shape[0], X_q.shape[2]
    # J(i,j) = dX_i/dxi_j = J_q[e](j,i,q)  (F/domain_integral_kernels.hpp:113-119)
    Jm = [[J_q[:, j, i, :] for j in range(dim)] for i in range(dim)]
    invJ = minv(Jm)
    dv = mdet(Jm)
    args = []
    for k, (uE, p) in enumerate(zip(trials_E, orders)):
        val, grad = interpolate(uE, p, q, dim, geom)  # parent space
        c = val.shape[1]
        vals, grads = _tensor_lists(val, grad)
        if k == wrt:
            # promote_each_to_dual_when BEFORE parent_to_physical (:181-186): seeds are value comps, then parent grads
            N = c + c * dim
            z = np.zeros((ne, nq, N))
            for a in range(c):
                g = z.copy()
                g[..., a] = 1.0
                vals[a] = Dual(vals[a], g)
                for j in range(dim):
                    g = z.copy()
                    g[..., c + a * dim + j] = 1.0
                    grads[a][j] = Dual(grads[a][j], g)
        grads = mdot(grads, invJ)  # parent_to_physical (F/finite_element.hpp:262-264)
        args.append((vals, grads))
    x = ([X_q[:, i, :] for i in range(dim)], Jm)
    source, flux = qf(0.0, x, state, *args)
    # physical_to_parent (F/finite_element.hpp:301-303): source *= dv ; flux = flux . inv(J^T) * dv
    ct = len(flux)
    source = [s * dv for s in source]
    flux = mscale(dv, mdot(flux, mT(invJ)))
    zero = np.zeros((ne, nq))
    src_v = np.stack([np.broadcast_to(value_of(s), (ne, nq)) for s in source], axis=1)
    flux_v = np.stack([np.stack([np.broadcast_to(value_of(flux[a][j]), (ne, nq)) for j in range(dim)], axis=1)
                       for a in range(ct)], axis=1)
    r_E = integrate(src_v, flux_v, test_order, q, dim, geom)
    if wrt < 0:
        return r_E, None
    ck = trials_E[wrt].shape[1]

    def partials(s):
        return s.g if isinstance(s, Dual) else np.zeros((ne, nq, ck + ck * dim))

    ds = np.stack([partials(s) for s in source], axis=2)  # [ne][nq][ct][N]
    df = np.stack([np.stack([partials(flux[a][j]) for j in range(dim)], axis=2) for a in range(ct)], axis=2)
    blocks = dict(d00=ds[..., :ck], d01=ds[..., ck:].reshape(ne, nq, ct, ck, dim), d10=df[..., :ck],
                  d11=df[..., ck:].reshape(ne, nq, ct, dim, ck, dim))
    return r_E, blocks


# ----------------------------------------------------------------------------------------------------------------------
# boundary integrals (F/boundary_integral_kernels.hpp:111-205): faces of a 3D mesh / segments of a 2D mesh, trial gradients stay in xi coordinates
# ----------------------------------------------------------------------------------------------------------------------
def vcross(A):
    """cross product of the two columns of a 3 x 2 matrix given as a list of rows; for a 2 x 1 matrix (the tangent of a
    segment in the plane) the vector perpendicular to its column, {A10, -A00} (F/tensor.hpp:906-928)"""
    if len(A) == 2:
        return [A[1][0], -1.0 * A[0][0]]
    a, b = [A[i][0] for i in range(3)], [A[i][1] for i in range(3)]
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def vnorm(v):
    s = v[0] * v[0]
    for c in v[1:]:
        s = s + c * c
    return d_sqrt(s)


def shape_aware_boundary(qf):
    """ShapeAwareBoundaryIntegrandWrapper (F/shape_aware_functional.hpp:474-499): integrand(t, x + p, args...) times the
    area correction |d(X+p)/dxi_1 x d(X+p)/dxi_2| / |dX/dxi_1 x dX/dxi_2| (:155-174)"""

    def wrapped(t, x, shape, *args):
        X, dX = x
        xp = ([X[i] + shape[0][i] for i in range(len(X))], madd(dX, shape[1]))
        corr = vnorm(vcross(xp[1])) / vnorm(vcross(dX))
        return [s * corr for s in qf(t, xp, *args)]

    return wrapped


def pressure_qf(pressure, geom_nonlin):
    """SolidMechanics::setPressure (src/serac/physics/solid_mechanics.hpp:1075-1101) with a constant pressure"""

    def qf(t, x, displacement, acceleration):
        X, dX = x
        dx = madd(dX, displacement[1]) if geom_nonlin else madd(dX, mscale(0.0, displacement[1]))
        n = vcross(dx)
        ref_area = vnorm(vcross(dX))
        return [pressure * (c / ref_area) for c in n]

    return qf


def traction_qf(traction):
    """SolidMechanics::setTraction (solid_mechanics.hpp:1030-1036) with a constant traction vector"""

    def qf(t, x, displacement, acceleration):
        return [-1.0 * tc + 0.0 * x[0][0] for tc in traction]

    return qf


def make_boundary_qf(kind, params):
    if kind == "solid_pressure":
        return shape_aware_boundary(pressure_qf(params[0], params[1] != 0.0))
    if kind == "solid_traction":
        return shape_aware_boundary(traction_qf(list(params)))
    raise KeyError(kind)


def _face_weights(geom, q):
    if is_simplex(geom):
        return simplex_quadrature(geom, q)[1]
    w = basis_tables(1, q)[2]
    return w if geom == SEGMENT else np.multiply.outer(w, w).ravel()


def boundary_evaluation_kernel(qf, trials_E, orders, X_q, J_q, test_order, q, face_geom, wrt=-1):
    """boundary_integral::evaluation_kernel_impl (F/boundary_integral_kernels.hpp:159-205) with batch_apply_qf (:134-156):
    X_q [nf][3][nq], J_q [nf][j][i][nq] = dX_i/dxi_j.  Returns r_E [nf][ct][ndof] and, when wrt >= 0, the derivative blocks
    d00 [nf][nq][ct][ck] and d01 [nf][nq][ct][ck][2] of the weighted source w.r.t. the value and the xi-gradient."""
    nf, nq = X_q.shape[0], X_q.shape[2]
    sdim, fd = X_q.shape[1], J_q.shape[1]  # space dimension, face dimension (2: faces of 3D meshes, 1: segments)
    dX = [[J_q[:, j, i, :] for j in range(fd)] for i in range(sdim)]  # sdim x fd
    args = []
    for k, (uE, p) in enumerate(zip(trials_E, orders)):
        val, grad = interpolate(uE, p, q, fd, face_geom)
        c = val.shape[1]
        vals, grads = _tensor_lists(val, grad)
        if k == wrt:
            N = c + c * fd
            z = np.zeros((nf, nq, N))
            for a in range(c):
                g = z.copy()
                g[..., a] = 1.0
                vals[a] = Dual(vals[a], g)
                for j in range(fd):
                    g = z.copy()
                    g[..., c + a * fd + j] = 1.0
                    grads[a][j] = Dual(grads[a][j], g)
        args.append((vals, grads))
    x = ([X_q[:, i, :] for i in range(sdim)], dX)
    source = qf(0.0, x, *args)
    measure = vnorm(vcross(dX))  # |dX/dxi_1 x dX/dxi_2| (:148-152)
    source = [s * measure for s in source]
    ct = len(source)
    src_v = np.stack([np.broadcast_to(value_of(s), (nf, nq)) for s in source], axis=1)
    r_E = integrate(src_v, None, test_order, q, fd, face_geom)
    if wrt < 0:
        return r_E, None
    ck = trials_E[wrt].shape[1]
    ds = np.stack([s.g if isinstance(s, Dual) else np.zeros((nf, nq, ck + ck * fd)) for s in source], axis=2)
    blocks = dict(d00=ds[..., :ck], d01=ds[..., ck:].reshape(nf, nq, ct, ck, fd))
    return r_E, blocks


def boundary_action_of_gradient_kernel(blocks, du_E, trial_order, test_order, q, face_geom):
    """boundary action_of_gradient_kernel (F/boundary_integral_kernels.hpp:252-277)"""
    fd = blocks["d01"].shape[-1]
    val, grad = interpolate(du_E, trial_order, q, fd, face_geom)
    v = np.transpose(val, (0, 2, 1))
    g = np.transpose(grad, (0, 3, 1, 2))
    src = np.einsum("eqab,eqb->eqa", blocks["d00"], v) + np.einsum("eqabl,eqbl->eqa", blocks["d01"], g)
    return integrate(np.transpose(src, (0, 2, 1)), None, test_order, q, fd, face_geom)


def action_of_gradient_kernel(blocks, du_E, trial_order, test_order, q, dim, geom=None):
    """action_of_gradient_kernel (F/domain_integral_kernels.hpp:277-304) with chain_rule (:223-237)"""
    val, grad = interpolate(du_E, trial_order, q, dim, geom)  # parent space, no J needed
    v = np.transpose(val, (0, 2, 1))  # [ne][nq][ck]
    g = np.transpose(grad, (0, 3, 1, 2))  # [ne][nq][ck][dim]
    src = np.einsum("eqab,eqb->eqa", blocks["d00"], v) + np.einsum("eqabl,eqbl->eqa", blocks["d01"], g)
    flx = np.einsum("eqakb,eqb->eqak", blocks["d10"], v) + np.einsum("eqakbl,eqbl->eqak", blocks["d11"], g)
    return integrate(np.transpose(src, (0, 2, 1)), np.transpose(flx, (0, 2, 3, 1)), test_order, q, dim, geom)


def element_gradient_kernel(blocks, trial_order, test_order, q, dim, geom=None):
    """element_gradient_kernel (F/domain_integral_kernels.hpp:327-361) with batch_apply_shape_fn
    (F/detail/hexahedron_H1.inl:132-167): K_e [ne][ck*ndk + J][ct*ndt + I] (trial-major, F/functional.hpp:599-601)"""
    d00, d01, d10, d11 = blocks["d00"], blocks["d01"], blocks["d10"], blocks["d11"]
    ne, nq, ct, ck = d00.shape
    nk = trial_order + 1
    ndk, ndt = nk**dim, (test_order + 1)**dim
    if geom is not None and is_simplex(geom):
        ndk, ndt = len(simplex_lattice(geom, trial_order)), len(simplex_lattice(geom, test_order))
    # shape function J and its parent gradient at every quadrature point = interpolation of the unit E-vector e_J
    phi, dphi = interpolate(np.eye(ndk)[None], trial_order, q, dim, geom)
    phi, dphi = phi[0], dphi[0]  # [J][nq], [J][dim][nq]
    K = np.zeros((ne, ck * ndk, ct * ndt))
    for b in range(ck):
        src = np.einsum("eqa,Jq->eJaq", d00[:, :, :, b], phi) + np.einsum("eqal,Jlq->eJaq", d01[:, :, :, b, :], dphi)
        flx = (np.einsum("eqak,Jq->eJakq", d10[:, :, :, :, b], phi) +
               np.einsum("eqakl,Jlq->eJakq", d11[:, :, :, :, b, :], dphi))
        rr = integrate(src.reshape(ne * ndk, ct, nq), flx.reshape(ne * ndk, ct, dim, nq), test_order, q, dim, geom)
        K[:, b * ndk:(b + 1) * ndk, :] = rr.reshape(ne, ndk, ct * ndt)
    return K


# ----------------------------------------------------------------------------------------------------------------------
# sparsity pattern + assembly (F/dof_numbering.hpp:290-351, F/functional.hpp:610-644)
# ----------------------------------------------------------------------------------------------------------------------
def element_vdofs(table, components, num_nodes, byvdim=True):
    """GetElementVDofs (F/element_restriction.cpp:434-441): [ne][c*npe], component-major"""
    table = np.asarray(table).astype(np.int64)
    return np.concatenate([vdof(table, c, components, num_nodes, byvdim) for c in range(components)], axis=1)


def csr_pattern(test_vdofs, trial_vdofs, num_rows):
    """GradientAssemblyLookupTables::init: sorted unique (row, col) pairs -> row_ptr, col_ind"""
    rows = np.repeat(test_vdofs[:, :, None], trial_vdofs.shape[1], axis=2).ravel()
    cols = np.repeat(trial_vdofs[:, None, :], test_vdofs.shape[1], axis=1).ravel()
    keys = np.unique(rows.astype(np.int64) << 32 | cols.astype(np.int64))
    r, c = keys >> 32, keys & 0xFFFFFFFF
    row_ptr = np.zeros(num_rows + 1, dtype=np.int32)
    np.add.at(row_ptr, r + 1, 1)
    return np.cumsum(row_ptr).astype(np.int32), c.astype(np.int32), keys


def csr_assemble(K_e, test_vdofs, trial_vdofs, keys):
    """values[LUT(row, col)] += K_e(e, i, j) in (e, i, j) order (F/functional.hpp:620-640)"""
    ne, ncol, nrow = K_e.shape
    rows = np.broadcast_to(test_vdofs[:, None, :], (ne, ncol, nrow)).ravel().astype(np.int64)
    cols = np.broadcast_to(trial_vdofs[:, :, None], (ne, ncol, nrow)).ravel().astype(np.int64)
    nz = np.searchsorted(keys, rows << 32 | cols)
    values = np.zeros(len(keys))
    np.add.at(values, nz, K_e.ravel())
    return values


# ----------------------------------------------------------------------------------------------------------------------
# named integrands (same names / parameters as the product's registry and the reference-oracle kinds)
# ----------------------------------------------------------------------------------------------------------------------
def parameterized_diffusion_qf(c0):
    """the test integrand of the L2-parameter path: H1<p>(H1<p> u, L2<0> k) -> source (c0 + k) u, flux k^2 grad u"""

    def qf(t, x, state, u, k):
        kv = k[0][0]
        return [(c0 + kv) * u[0][0]], [[(kv * kv) * g for g in u[1][0]]]

    return qf


def make_qf(kind, params):
    if kind == "param_diffusion":
        return parameterized_diffusion_qf(params[0])
    if kind == "solid_neohookean":
        return shape_aware(material_stress_qf(NeoHookean(*params[:3]), params[3] != 0.0))
    if kind == "solid_linear_isotropic":
        return shape_aware(material_stress_qf(LinearIsotropic(*params[:3]), params[3] != 0.0))
    if kind == "solid_j2smallstrain_linear":
        return shape_aware(material_stress_qf(J2SmallStrainLinear(*params[:6]), params[6] != 0.0))
    if kind == "solid_j2smallstrain_powerlaw":  # E nu sigma_y n eps0 Hk density gn (solid_material.hpp:154-173)
        E, nu, sy, n, e0, Hk, rho = params[:7]
        law = ((lambda e: sy * (1.0 + e / e0)**(1.0 / n)), (lambda e: sy / (n * e0) * (1.0 + e / e0)**(1.0 / n - 1.0)))
        return shape_aware(material_stress_qf(J2SmallStrainLinear(E, nu, sy, 0.0, Hk, rho, hardening=law), params[7] != 0.0))
    if kind == "solid_j2smallstrain_voce":  # E nu sigma_y sigma_sat strain_constant Hk density gn (:180-198)
        E, nu, sy, ss, c, Hk, rho = params[:7]
        law = ((lambda e: ss - (ss - sy) * np.exp(-e / c)), (lambda e: (ss - sy) / c * np.exp(-e / c)))
        return shape_aware(material_stress_qf(J2SmallStrainLinear(E, nu, sy, 0.0, Hk, rho, hardening=law), params[7] != 0.0))
    if kind == "solid_stvk":
        return shape_aware(material_stress_qf(StVenantKirchhoff(*params[:3]), params[3] != 0.0))
    if kind == "solid_j2_linear":
        return shape_aware(material_stress_qf(J2FiniteLinear(*params[:5]), params[5] != 0.0))
    if kind == "thermal_linear_conductor":
        return shape_aware(thermal_qf(LinearIsotropicConductor(*params[:3])))
    if kind == "thermoelastic_solid":
        return shape_aware(thermoelastic_solid_qf(GreenSaintVenantThermoelastic(*params[:7]), params[7] != 0.0))
    if kind == "thermoelastic_thermal":
        return shape_aware(thermoelastic_thermal_qf(GreenSaintVenantThermoelastic(*params[:7])))
    raise KeyError(kind)
