// serac_b200: a named quantity of interest -- `Functional<double(H1<p,dim>)>` (src/serac/numerics/functional/
// functional_qoi.inl) with the integrand  f(x, u) = a (1 + x_0) u.u + b/2 |grad u|^2 + sin(x_1):  depends on the
// position, the value and the gradient, so every derivative block of the QOI path is exercised.  The physics modules
// build their QOIs from user functors through the same make_qoi_kernel_table (kernel_table.cuh).
#include <cmath>

#include "registry.hpp"

namespace sb {
namespace {

#ifndef SB_REG_GEOM
#error "compile with -DSB_REG_GEOM=CUBE|TETRAHEDRON -DSB_REG_P=1|2 -DSB_REG_TAG=hex_p2"
#endif
#define SB_STR2(x) #x
#define SB_STR(x) SB_STR2(x)

constexpr Geometry g   = Geometry::SB_REG_GEOM;
constexpr int      p   = SB_REG_P;
constexpr int      dim = dimension_of(g);

#ifndef SB_REG_BOUNDARY
struct QuadraticQoi {
  double a, b;
  template <typename X, typename U>
  SERAC_HOST_DEVICE auto operator()(double, X x, U u) const
  {
    auto uv = get<0>(u);
    auto du = get<1>(u);
    auto X0 = get<0>(x);
    return a * (1.0 + X0[0]) * dot(uv, uv) + (0.5 * b) * inner(du, du) + ::sin(X0[1]);
  }
};

#endif
#ifdef SB_REG_BOUNDARY
// boundary variant: g is the FACE geometry; f(x, u) = a (1 + x_0) u.u + b/2 |du/dxi|^2 (trial gradients stay in the
// face's parent coordinates on a boundary, boundary_integral_kernels.hpp:134-156), weighted by the surface element
constexpr int sdim = dim + 1;
struct QuadraticBoundaryQoi {
  double a, b;
  template <typename X, typename U>
  SERAC_HOST_DEVICE auto operator()(double, X x, U u) const
  {
    auto uv = get<0>(u);
    auto du = get<1>(u);
    auto X0 = get<0>(x);
    return a * (1.0 + X0[0]) * dot(uv, uv) + (0.5 * b) * inner(du, du);
  }
};
using BSig = BoundarySignature<g, p + 1, QOI, H1<p, sdim>>;
long make_bqoi(const double* q, int n, void* out, size_t cap)
{
  if (n != 2) return -1;
  return store_blob(QuadraticBoundaryQoi{q[0], q[1]}, out, cap);
}
static Registrar r_bqoi({"qoi_boundary_quadratic_" SB_STR(SB_REG_TAG), make_boundary_kernel_table<BSig, QuadraticBoundaryQoi>(),
                         &make_bqoi, "{a, b}: boundary integrand a (1 + x_0) u.u + b/2 |du/dxi|^2"});
#ifdef SB_REG_TAG2
static Registrar r_bqoi2({"qoi_boundary_quadratic_" SB_STR(SB_REG_TAG2), make_boundary_kernel_table<BSig, QuadraticBoundaryQoi>(),
                          &make_bqoi, "{a, b}: boundary integrand a (1 + x_0) u.u + b/2 |du/dxi|^2"});
#endif
#else
using Sig = Signature<g, p + 1, QOI, H1<p, dim>>;

long make_qoi(const double* q, int n, void* out, size_t cap)
{
  if (n != 2) return -1;
  return store_blob(QuadraticQoi{q[0], q[1]}, out, cap);
}

static Registrar r_qoi({"qoi_quadratic_" SB_STR(SB_REG_TAG), make_qoi_kernel_table<Sig, QuadraticQoi>(), &make_qoi,
                        "{a, b}: integrand a (1 + x_0) u.u + b/2 |grad u|^2 + sin(x_1)"});
#endif

}  // namespace
}  // namespace sb
