// serac_b200: named heat-transfer integrands -- what HeatTransfer::setMaterial registers with its Functional
// (src/serac/physics/heat_transfer.hpp:381-437): ShapeAwareIntegrandWrapper<ThermalMaterialIntegrand<Material>>
// on H1<p>(H1<1,dim> shape, H1<p> temperature, H1<p> dT/dt).
#include "registry.hpp"

namespace sb {
namespace {

using heat_transfer::LinearIsotropicConductor;

template <typename S, typename Material>
long make_thermal(const double* p, int n, void* out, size_t cap)
{
  if (n != 3) return -1;
  typename S::template qf<Material> qf{ThermalMaterialIntegrand<Material>(Material(p[0], p[1], p[2]))};
  return store_blob(qf, out, cap);
}

#ifndef SB_REG_GEOM
#error "compile with -DSB_REG_GEOM=CUBE|SQUARE -DSB_REG_P=1|2|3 -DSB_REG_TAG=hex_p2"
#endif
#define SB_STR2(x) #x
#define SB_STR(x) SB_STR2(x)

using S = ThermalSig<Geometry::SB_REG_GEOM, SB_REG_P>;

static Registrar r_lin({"thermal_linear_conductor_" SB_STR(SB_REG_TAG),
                        make_kernel_table<S::type, S::qf<LinearIsotropicConductor>, Nothing>(),
                        &make_thermal<S, LinearIsotropicConductor>, "{density, specific_heat_capacity, conductivity}"});

}  // namespace
}  // namespace sb
