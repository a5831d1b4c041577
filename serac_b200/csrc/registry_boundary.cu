// serac_b200: named boundary integrands -- what SolidMechanics::setTraction / setPressure register with their Functional
// (src/serac/physics/solid_mechanics.hpp:1024-1101), wrapped by ShapeAwareFunctional::AddBoundaryIntegral
// (shape_aware_functional.hpp:474-520), on the boundary faces of H1<p,dim>(H1<1,dim> shape, H1<p,dim> u, H1<p,dim> a):
// quadrilaterals (hex meshes), triangles (tet meshes) and segments (2-D meshes, F/detail/segment_H1.inl).
#include "registry.hpp"

namespace sb {
namespace {

#ifndef SB_REG_GEOM
#error "compile with -DSB_REG_GEOM=SQUARE -DSB_REG_P=1|2|3 -DSB_REG_TAG=hex_p2   (geometry = the FACE geometry)"
#endif
#define SB_STR2(x) #x
#define SB_STR(x) SB_STR2(x)

constexpr Geometry face = Geometry::SB_REG_GEOM;
constexpr int      p    = SB_REG_P;
constexpr int      sdim = dimension_of(face) + 1;
using Sig       = BoundarySignature<face, p + 1, H1<p, sdim>, H1<1, sdim>, H1<p, sdim>, H1<p, sdim>>;
using TractionQF = ShapeAwareBoundaryIntegrandWrapper<TractionIntegrand<UniformTraction<sdim>>>;
using PressureQF = ShapeAwareBoundaryIntegrandWrapper<PressureIntegrand<UniformPressure>>;

long make_traction(const double* q, int n, void* out, size_t cap)
{
  if (n != sdim) return -1;
  UniformTraction<sdim> f{};
  for (int i = 0; i < sdim; i++) f.traction[i] = q[i];
  TractionQF qf{TractionIntegrand<UniformTraction<sdim>>(f)};
  return store_blob(qf, out, cap);
}

long make_pressure(const double* q, int n, void* out, size_t cap)
{
  if (n != 2) return -1;
  auto       gn = (q[1] != 0.0) ? GeometricNonlinearities::On : GeometricNonlinearities::Off;
  PressureQF qf{PressureIntegrand<UniformPressure>(UniformPressure{q[0]}, gn)};
  return store_blob(qf, out, cap);
}

static Registrar r_traction({"solid_traction_" SB_STR(SB_REG_TAG), make_boundary_kernel_table<Sig, TractionQF>(),
                             &make_traction, "{traction_x, traction_y[, traction_z]}"});
static Registrar r_pressure({"solid_pressure_" SB_STR(SB_REG_TAG), make_boundary_kernel_table<Sig, PressureQF>(),
                             &make_pressure, "{pressure, geometric_nonlinearity(0|1)}"});
#ifdef SB_REG_TAG2  // segment faces bound quadrilateral and triangular meshes alike: the same tables under both names
static Registrar r_traction2({"solid_traction_" SB_STR(SB_REG_TAG2), make_boundary_kernel_table<Sig, TractionQF>(),
                              &make_traction, "{traction_x, traction_y[, traction_z]}"});
static Registrar r_pressure2({"solid_pressure_" SB_STR(SB_REG_TAG2), make_boundary_kernel_table<Sig, PressureQF>(),
                              &make_pressure, "{pressure, geometric_nonlinearity(0|1)}"});
#endif

}  // namespace
}  // namespace sb
