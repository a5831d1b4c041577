// serac_b200: function-space tags, geometry ids and the 1-D basis / quadrature tables of the tensor-product
// H1 elements.
//
// Mirrors (API names and numerical content, not code) of
//   src/serac/numerics/functional/finite_element.hpp:190-240   (Family, H1<p,c>, L2<p,c>, QOI)
//   src/serac/numerics/functional/geometry.hpp:25-60           (num_quadrature_points, dimension_of)
//   src/serac/numerics/functional/polynomials.hpp:42-55,271-283 (Gauss-Legendre nodes / weights on [0,1])
//   src/serac/numerics/functional/polynomials.hpp:493-540      (Gauss-Lobatto Lagrange bases, n <= 4)
//   src/serac/numerics/functional/detail/hexahedron_H1.inl:96-130 (B(q,d), G(q,d))
// The node / weight literals are the reference's own 15-16 digit values (SURVEY.md §8a parity note): they differ
// from the exact irrationals at 1e-15, which is inside the 1e-12 budget but costs nothing to match.
#pragma once

#include <cmath>
#include <utility>

#include "serac_b200/tensor.hpp"

namespace serac {

enum class Family
{
  QOI,
  H1,
  HCURL,
  HDIV,
  L2
};

template <int p, int c = 1>
struct H1 {
  static constexpr int    order      = p;
  static constexpr int    components = c;
  static constexpr Family family     = Family::H1;
};
template <int p, int c = 1>
struct L2 {
  static constexpr int    order      = p;
  static constexpr int    components = c;
  static constexpr Family family     = Family::L2;
};
struct QOI {
  static constexpr int    order      = 0;
  static constexpr int    components = 1;
  static constexpr Family family     = Family::QOI;
};

/// geometry ids share mfem::Geometry::Type's numeric values so tables from an mfem host can be passed through
enum class Geometry : int
{
  POINT       = 0,
  SEGMENT     = 1,
  TRIANGLE    = 2,
  SQUARE      = 3,
  TETRAHEDRON = 4,
  CUBE        = 5
};

SB_HD int dimension_of(Geometry g)
{
  switch (g) {
    case Geometry::SEGMENT:
      return 1;
    case Geometry::TRIANGLE:
    case Geometry::SQUARE:
      return 2;
    case Geometry::TETRAHEDRON:
    case Geometry::CUBE:
      return 3;
    default:
      return 0;
  }
}

SB_HD int num_quadrature_points(Geometry g, int Q)
{
  switch (g) {
    case Geometry::SEGMENT:
      return Q;
    case Geometry::TRIANGLE:
      return (Q * (Q + 1)) / 2;
    case Geometry::SQUARE:
      return Q * Q;
    case Geometry::TETRAHEDRON:
      return (Q * (Q + 1) * (Q + 2)) / 6;
    case Geometry::CUBE:
      return Q * Q * Q;
    default:
      return -1;
  }
}

SB_HD int ipow(int b, int e)
{
  int r = 1;
  for (int i = 0; i < e; i++) r *= b;
  return r;
}

namespace b200 {

constexpr int MAX_Q = 4;  ///< p <= 3 (polynomials.hpp:496) so Q = p + 1 <= 4 (functional.hpp:193)

/// Gauss-Legendre nodes on [0,1], ascending
inline void gauss_legendre_nodes(int n, double* x)
{
  static const double t1[] = {0.50000000000000000};
  static const double t2[] = {0.2113248654051871, 0.7886751345948129};
  static const double t3[] = {0.1127016653792583, 0.500000000000000, 0.887298334620742};
  static const double t4[] = {0.0694318442029737, 0.330009478207572, 0.669990521792428, 0.930568155797026};
  const double*       t[]  = {nullptr, t1, t2, t3, t4};
  for (int i = 0; i < n; i++) x[i] = t[n][i];
}

/// Gauss-Legendre weights on [0,1]
inline void gauss_legendre_weights(int n, double* w)
{
  static const double t1[] = {1.000000000000000};
  static const double t2[] = {0.500000000000000, 0.500000000000000};
  static const double t3[] = {0.277777777777778, 0.444444444444444, 0.277777777777778};
  static const double t4[] = {0.173927422568727, 0.326072577431273, 0.326072577431273, 0.173927422568727};
  const double*       t[]  = {nullptr, t1, t2, t3, t4};
  for (int i = 0; i < n; i++) w[i] = t[n][i];
}

/// Lagrange polynomials through the n Gauss-Lobatto points of [0,1], evaluated at x
inline void gauss_lobatto_interpolation(int n, double x, double* N)
{
  constexpr double sqrt5 = 2.23606797749978981;
  switch (n) {
    case 1:
      N[0] = 1.0;
      break;
    case 2:
      N[0] = 1.0 - x;
      N[1] = x;
      break;
    case 3:
      N[0] = (-1.0 + x) * (-1.0 + 2.0 * x);
      N[1] = -4.0 * (-1.0 + x) * x;
      N[2] = x * (-1.0 + 2.0 * x);
      break;
    case 4:
      N[0] = -(-1.0 + x) * (1.0 + 5.0 * (-1.0 + x) * x);
      N[1] = -0.5 * sqrt5 * (5.0 + sqrt5 - 10.0 * x) * (-1.0 + x) * x;
      N[2] = -0.5 * sqrt5 * (-1.0 + x) * x * (-5.0 + sqrt5 + 10.0 * x);
      N[3] = x * (1.0 + 5.0 * (-1.0 + x) * x);
      break;
  }
}

/// derivatives of the above
inline void gauss_lobatto_interpolation_derivative(int n, double x, double* dN)
{
  constexpr double sqrt5 = 2.23606797749978981;
  switch (n) {
    case 1:
      dN[0] = 0.0;
      break;
    case 2:
      dN[0] = -1.0;
      dN[1] = 1.0;
      break;
    case 3:
      dN[0] = -3.0 + 4.0 * x;
      dN[1] = 4.0 - 8.0 * x;
      dN[2] = -1.0 + 4.0 * x;
      break;
    case 4:
      dN[0] = -6.0 + 5.0 * (4.0 - 3.0 * x) * x;
      dN[1] = 2.5 * (1.0 + sqrt5 + 2.0 * x * (-1.0 - 3.0 * sqrt5 + 3.0 * sqrt5 * x));
      dN[2] = -2.5 * (-1.0 + sqrt5 + 2.0 * x * (1.0 - 3.0 * sqrt5 + 3.0 * sqrt5 * x));
      dN[3] = 1.0 + 5.0 * x * (-2.0 + 3.0 * x);
      break;
  }
}

/// 1-D tables of one H1 space of order p sampled at the q-point Gauss-Legendre rule.
/// B[qi*n + d] = d-th nodal polynomial at the qi-th point, G its derivative, w the quadrature weights.
struct BasisTable1D {
  int    n;  ///< nodes per direction (p + 1)
  int    q;  ///< quadrature points per direction
  double B[MAX_Q * MAX_Q];
  double G[MAX_Q * MAX_Q];
  double w[MAX_Q];
};

inline BasisTable1D make_basis_table(int p, int q)
{
  BasisTable1D t{};
  t.n = p + 1;
  t.q = q;
  double x[MAX_Q];
  gauss_legendre_nodes(q, x);
  gauss_legendre_weights(q, t.w);
  for (int i = 0; i < q; i++) {
    gauss_lobatto_interpolation(t.n, x[i], &t.B[i * t.n]);
    gauss_lobatto_interpolation_derivative(t.n, x[i], &t.G[i * t.n]);
  }
  return t;
}


// ------------------------------------------------------------------------------------------------
// simplex H1 elements (triangle / tetrahedron), p <= 3
//   src/serac/numerics/functional/detail/triangle_H1.inl, tetrahedron_H1.inl   (nodal bases, dense loops)
//   src/serac/numerics/functional/polynomials.hpp:57-260, 285-365               (simplex quadrature rules)
// The reference spells every nodal polynomial out; here the same (unique) Lagrange basis of P_p through the same
// nodes -- lexicographic (x fastest), Gauss-Lobatto points along edges, centroids inside faces -- is obtained by
// inverting the monomial Vandermonde matrix once on the host in long double.  Quadrature points and weights are the
// reference's literal tables (same order, so stored per-point derivatives line up).
// ------------------------------------------------------------------------------------------------
SB_HD bool is_simplex(Geometry g) { return g == Geometry::TRIANGLE || g == Geometry::TETRAHEDRON; }

/// geometries whose kernels interpolate / integrate by dense sums against a {phi, dphi/dxi} table: the simplices and the
/// segment (the boundary faces of 2-D meshes, F/detail/segment_H1.inl: with one direction there is nothing to factorise)
SB_HD bool uses_dense_table(Geometry g) { return is_simplex(g) || g == Geometry::SEGMENT; }

/// H1 nodes per element
SB_HD int num_nodes(Geometry g, int p)
{
  switch (g) {
    case Geometry::SEGMENT:
      return p + 1;
    case Geometry::TRIANGLE:
      return (p + 1) * (p + 2) / 2;  // p = 0 -> 1: the per-element total of a quantity of interest
    case Geometry::SQUARE:
      return (p + 1) * (p + 1);
    case Geometry::TETRAHEDRON:
      return (p + 1) * (p + 2) * (p + 3) / 6;
    case Geometry::CUBE:
      return (p + 1) * (p + 1) * (p + 1);
    default:
      return 0;
  }
}

constexpr int MAX_SIMPLEX_NODES = 20;  ///< tetrahedron p = 3
constexpr int MAX_SIMPLEX_QPTS  = 20;  ///< tetrahedron Q = 4

/// quadrature rule of a simplex: returns the number of points; xi[i*dim + d], w[i]
inline int simplex_quadrature(Geometry g, int Q, double* xi, double* w)
{
  if (g == Geometry::TRIANGLE) {
    static const double x1[] = {0.333333333333333, 0.333333333333333};
    static const double w1[] = {0.5};
    static const double x2[] = {0.166666666666667, 0.166666666666667, 0.166666666666667, 0.666666666666667,
                                0.666666666666667, 0.166666666666667};
    static const double w2[] = {0.166666666666667, 0.166666666666667, 0.166666666666667};
    static const double x3[] = {0.09157621350978,  0.09157621350978,  0.09157621350978,  0.81684757298044,
                                0.81684757298044,  0.09157621350978,  0.108103018168071, 0.445948490915964,
                                0.445948490915964, 0.108103018168071, 0.445948490915964, 0.445948490915964};
    static const double w3[] = {0.0549758718276665, 0.0549758718276665, 0.0549758718276665,
                                0.111690794839,     0.111690794839,     0.111690794839};
    static const double x4[] = {0.055564052669793, 0.055564052669793, 0.055564052669793, 0.888871894660413,
                                0.888871894660413, 0.055564052669793, 0.070255540518384, 0.634210747745723,
                                0.634210747745723, 0.070255540518384, 0.634210747745723, 0.295533711735893,
                                0.070255540518384, 0.295533711735893, 0.295533711735893, 0.070255540518384,
                                0.295533711735893, 0.634210747745723, 0.333333333333333, 0.333333333333333};
    static const double w4[] = {0.0209777564983245, 0.0209777564983245, 0.0209777564983245, 0.0560492060354435,
                                0.0560492060354435, 0.0560492060354435, 0.0560492060354435, 0.0560492060354435,
                                0.0560492060354435, 0.100771494292365};
    const double* X[] = {nullptr, x1, x2, x3, x4};
    const double* W[] = {nullptr, w1, w2, w3, w4};
    if (Q < 1 || Q > 4) return -1;
    const int n = Q * (Q + 1) / 2;
    for (int i = 0; i < n; i++) {
      w[i] = W[Q][i];
      for (int d = 0; d < 2; d++) xi[i * 2 + d] = X[Q][i * 2 + d];
    }
    return n;
  }
  if (g == Geometry::TETRAHEDRON) {
    constexpr double a2 = 0.585410196624967936, b2 = 0.138196601125011008;
    constexpr double a3 = 0.77849529482132992, b3 = 0.0738349017262233984, c3 = 0.406244343884051008,
                     d3 = 0.0937556561159491072;
    constexpr double a4 = 0.902942215818267904, b4 = 0.0323525947272438976, c4 = 0.262682583887778976,
                     d4 = 0.616596533061936896, e4 = 0.0603604415251420928, f4 = 0.309769304272861952,
                     g4 = 0.0706920871814129024;
    static const double x1[] = {0.25, 0.25, 0.25};
    static const double w1[] = {0.166666666666666656};
    static const double x2[] = {a2, b2, b2, b2, a2, b2, b2, b2, a2, b2, b2, b2};
    static const double w2[] = {0.0416666666666666624, 0.0416666666666666624, 0.0416666666666666624,
                                0.0416666666666666624};
    static const double x3[] = {a3, b3, b3, b3, a3, b3, b3, b3, a3, b3, b3, b3, c3, c3, d3,
                                c3, d3, c3, c3, d3, d3, d3, c3, c3, d3, c3, d3, d3, d3, c3};
    constexpr double    u3   = 0.00793885580720148224, v3 = 0.022485207239643504;
    static const double w3[] = {u3, u3, u3, u3, v3, v3, v3, v3, v3, v3};
    static const double x4[] = {a4, b4, b4, b4, a4, b4, b4, b4, a4, b4, b4, b4, c4, d4, e4, d4, c4, e4, c4, e4,
                                d4, d4, e4, c4, c4, e4, e4, d4, e4, e4, e4, c4, d4, e4, d4, c4, e4, c4, e4, e4,
                                d4, e4, e4, e4, c4, e4, e4, d4, f4, f4, f4, f4, f4, g4, f4, g4, f4, g4, f4, f4};
    constexpr double    u4   = 0.00117784579907825008, v4 = 0.0078331114953146176, y4 = 0.0169894863816446688;
    static const double w4[] = {u4, u4, u4, u4, v4, v4, v4, v4, v4, v4, v4, v4, v4, v4, v4, v4, y4, y4, y4, y4};
    const double* X[] = {nullptr, x1, x2, x3, x4};
    const double* W[] = {nullptr, w1, w2, w3, w4};
    if (Q < 1 || Q > 4) return -1;
    const int n = Q * (Q + 1) * (Q + 2) / 6;
    for (int i = 0; i < n; i++) {
      w[i] = W[Q][i];
      for (int d = 0; d < 3; d++) xi[i * 3 + d] = X[Q][i * 3 + d];
    }
    return n;
  }
  return -1;
}

/// parent coordinates of the H1 nodes of a simplex of order p in the reference's (lexicographic) numbering
inline int simplex_nodes(Geometry g, int p, double* x)
{
  constexpr double sqrt5 = 2.23606797749978981;
  const double     gl2[] = {0.0, 1.0};
  const double     gl3[] = {0.0, 0.5, 1.0};
  const double     gl4[] = {0.0, 0.5 - 0.1 * sqrt5, 0.5 + 0.1 * sqrt5, 1.0};
  const double*    gl    = (p == 1) ? gl2 : (p == 2) ? gl3 : gl4;
  const int        dim   = dimension_of(g);
  int              count = 0;
  for (int k = 0; k <= (dim == 3 ? p : 0); k++) {
    for (int j = 0; j <= p - k; j++) {
      for (int i = 0; i <= p - k - j; i++) {
        const int l       = p - i - j - k;
        const int nonzero = (i > 0) + (j > 0) + (k > 0) + (l > 0);
        double    c[3];
        if (nonzero == 3) {  // the single interior node of a face at p = 3: its centroid
          c[0] = (i > 0) ? 1.0 / 3.0 : 0.0;
          c[1] = (j > 0) ? 1.0 / 3.0 : 0.0;
          c[2] = (k > 0) ? 1.0 / 3.0 : 0.0;
        } else {  // vertices and edge nodes: Gauss-Lobatto points along the edge
          c[0] = gl[i];
          c[1] = gl[j];
          c[2] = gl[k];
          // on an edge that does not touch the origin the two coordinates must sum to one
          if (nonzero == 2 && l == 0) {
            if (i > 0 && j > 0) c[1] = 1.0 - c[0];
            if (i > 0 && k > 0) c[2] = 1.0 - c[0];
            if (j > 0 && k > 0) c[2] = 1.0 - c[1];
          }
        }
        for (int d = 0; d < dim; d++) x[count * dim + d] = c[d];
        count++;
      }
    }
  }
  return count;
}

/// shape-function tables of one simplex H1 space at one quadrature rule:
/// tab[(qpt * ndof + k) * (1 + dim) + 0] = phi_k(xi_qpt), [.. + 1 + d] = d phi_k / d xi_d; then the nq weights.
struct SimplexTable {
  int    ndof, nq, dim;
  double tab[MAX_SIMPLEX_QPTS * MAX_SIMPLEX_NODES * 4 + MAX_SIMPLEX_QPTS];
  int    doubles() const { return nq * ndof * (1 + dim) + nq; }
  const double* weights() const { return tab + nq * ndof * (1 + dim); }
};

/// phi_k and its parent gradient at arbitrary points (pts[i*dim + d]); out[(i*ndof + k)*(1+dim) + ...]
inline int simplex_shape_functions(Geometry g, int p, int npts, const double* pts, double* out)
{
  const int dim = dimension_of(g);
  double    nodes[MAX_SIMPLEX_NODES * 3];
  const int nd = simplex_nodes(g, p, nodes);
  // monomials x^a y^b z^c with a + b + c <= p
  int ex[MAX_SIMPLEX_NODES][3];
  int nm = 0;
  for (int c = 0; c <= (dim == 3 ? p : 0); c++)
    for (int b = 0; b <= p - c; b++)
      for (int a = 0; a <= p - c - b; a++) {
        ex[nm][0] = a;
        ex[nm][1] = b;
        ex[nm][2] = c;
        nm++;
      }
  auto mono = [&](const long double* x, int m, int deriv) -> long double {
    long double v = 1.0L;
    for (int d = 0; d < dim; d++) {
      int e = ex[m][d];
      if (d == deriv) {
        if (e == 0) return 0.0L;
        v *= e;
        e -= 1;
      }
      for (int t = 0; t < e; t++) v *= x[d];
    }
    return v;
  };
  // V(n, m) = monomial m at node n; coefficients C = V^{-1} (column k = basis function k)
  long double V[MAX_SIMPLEX_NODES][2 * MAX_SIMPLEX_NODES];
  for (int n = 0; n < nd; n++) {
    long double x[3] = {nodes[n * dim], nodes[n * dim + 1], dim == 3 ? nodes[n * dim + 2] : 0.0L};
    for (int m = 0; m < nd; m++) {
      V[n][m]      = mono(x, m, -1);
      V[n][nd + m] = (n == m) ? 1.0L : 0.0L;
    }
  }
  for (int col = 0; col < nd; col++) {  // Gauss-Jordan with partial pivoting
    int piv = col;
    for (int r = col + 1; r < nd; r++)
      if (std::fabs((double)V[r][col]) > std::fabs((double)V[piv][col])) piv = r;
    for (int m = 0; m < 2 * nd; m++) std::swap(V[col][m], V[piv][m]);
    const long double inv = 1.0L / V[col][col];
    for (int m = 0; m < 2 * nd; m++) V[col][m] *= inv;
    for (int r = 0; r < nd; r++) {
      if (r == col) continue;
      const long double f = V[r][col];
      if (f != 0.0L)
        for (int m = 0; m < 2 * nd; m++) V[r][m] -= f * V[col][m];
    }
  }
  // phi_k(x) = sum_m C(m, k) mono_m(x), C(m, k) = V[m][nd + k]
  for (int i = 0; i < npts; i++) {
    long double x[3] = {pts[i * dim], pts[i * dim + 1], dim == 3 ? pts[i * dim + 2] : 0.0L};
    for (int k = 0; k < nd; k++) {
      for (int d = -1; d < dim; d++) {
        long double s = 0.0L;
        for (int m = 0; m < nd; m++) s += V[m][nd + k] * mono(x, m, d);
        out[(i * nd + k) * (1 + dim) + (d + 1)] = (double)s;
      }
    }
  }
  return nd;
}

inline SimplexTable make_simplex_table(Geometry g, int p, int Q)
{
  SimplexTable t{};
  t.dim = dimension_of(g);
  if (g == Geometry::SEGMENT) {  // segment_H1.inl:20-118: Gauss-Lobatto nodes (ascending), Gauss-Legendre rule on [0, 1]
    const BasisTable1D b = make_basis_table(p, Q);
    t.nq                 = Q;
    t.ndof               = p + 1;
    for (int i = 0; i < Q; i++) {
      for (int k = 0; k < t.ndof; k++) {
        t.tab[(i * t.ndof + k) * 2 + 0] = b.B[i * b.n + k];
        t.tab[(i * t.ndof + k) * 2 + 1] = b.G[i * b.n + k];
      }
      t.tab[Q * t.ndof * 2 + i] = b.w[i];
    }
    return t;
  }
  double xi[MAX_SIMPLEX_QPTS * 3], w[MAX_SIMPLEX_QPTS];
  t.nq   = simplex_quadrature(g, Q, xi, w);
  if (p == 0) {  // quantity of interest: one "shape function" = 1 (F/detail/qoi.inl:24)
    t.ndof = 1;
    for (int i = 0; i < t.nq; i++) {
      t.tab[i * (1 + t.dim)] = 1.0;
      for (int d = 0; d < t.dim; d++) t.tab[i * (1 + t.dim) + 1 + d] = 0.0;
    }
  } else {
    t.ndof = simplex_shape_functions(g, p, t.nq, xi, t.tab);
  }
  for (int i = 0; i < t.nq; i++) t.tab[t.nq * t.ndof * (1 + t.dim) + i] = w[i];
  return t;
}

}  // namespace b200
}  // namespace serac
