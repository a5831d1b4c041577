// serac_b200: named thermomechanics integrands -- what Thermomechanics::setMaterial registers with the two Functionals
// it owns (src/serac/physics/thermomechanics.hpp:376-384, 534-537):
//   SolidMechanics<p,3,Parameters<H1<p>>> :  H1<p,3>( shape H1<1,3>, u H1<p,3>, a H1<p,3>, temperature H1<p> )
//       ShapeAwareIntegrandWrapperWithState<MaterialStressFunctor<MechanicalMaterialInterface<Material>>>
//   HeatTransfer<p,3,Parameters<H1<p,3>>>  :  H1<p>  ( shape H1<1,3>, T H1<p>, dT/dt H1<p>, displacement H1<p,3> )
//       ShapeAwareIntegrandWrapper<ThermalMaterialIntegrand<ThermalMaterialInterface<Material>>>
// Both are differentiated with dual numbers end to end (any argument).
#include "registry.hpp"

namespace sb {
namespace {

#ifndef SB_REG_GEOM
#error "compile with -DSB_REG_GEOM=CUBE|TETRAHEDRON -DSB_REG_P=1|2 -DSB_REG_TAG=tet_p1"
#endif
#define SB_STR2(x) #x
#define SB_STR(x) SB_STR2(x)

constexpr Geometry g   = Geometry::SB_REG_GEOM;
constexpr int      p   = SB_REG_P;
constexpr int      dim = 3;
using Material         = GreenSaintVenantThermoelasticMaterial;
using Mech             = thermomechanics::MechanicalMaterialInterface<Material>;
using Therm            = thermomechanics::ThermalMaterialInterface<Material>;

using solid_test   = H1<p, dim>;
using SolidSigT    = Signature<g, p + 1, solid_test, H1<1, dim>, H1<p, dim>, H1<p, dim>, H1<p>>;
using SolidQF      = ShapeAwareIntegrandWrapperWithState<MaterialStressFunctor<Mech>, dim, solid_test, H1<p, dim>, H1<p, dim>, H1<p>>;
using thermal_test = H1<p>;
using ThermalSigT  = Signature<g, p + 1, thermal_test, H1<1, dim>, H1<p>, H1<p>, H1<p, dim>>;
using ThermalQF    = ShapeAwareIntegrandWrapper<ThermalMaterialIntegrand<Therm>, dim, thermal_test, H1<p>, H1<p>, H1<p, dim>>;

Material make_material(const double* q) { return Material{q[0], q[1], q[2], q[3], q[4], q[5], q[6]}; }

long make_solid(const double* q, int n, void* out, size_t cap)
{
  if (n != 8) return -1;
  auto    gn = (q[7] != 0.0) ? GeometricNonlinearities::On : GeometricNonlinearities::Off;
  SolidQF qf{MaterialStressFunctor<Mech>(Mech(make_material(q)), gn)};
  return store_blob(qf, out, cap);
}

long make_thermal(const double* q, int n, void* out, size_t cap)
{
  if (n != 7) return -1;
  ThermalQF qf{ThermalMaterialIntegrand<Therm>(Therm(make_material(q)))};
  return store_blob(qf, out, cap);
}

static Registrar r_solid({"thermoelastic_solid_" SB_STR(SB_REG_TAG), make_kernel_table<SolidSigT, SolidQF, Material::State>(),
                          &make_solid, "{density, E, nu, C_v, alpha, theta_ref, k, geometric_nonlinearity(0|1)}"});
static Registrar r_thermal({"thermoelastic_thermal_" SB_STR(SB_REG_TAG), make_kernel_table<ThermalSigT, ThermalQF, Nothing>(),
                            &make_thermal, "{density, E, nu, C_v, alpha, theta_ref, k}"});

}  // namespace
}  // namespace sb
