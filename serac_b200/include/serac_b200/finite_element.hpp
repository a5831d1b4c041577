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

}  // namespace b200
}  // namespace serac
