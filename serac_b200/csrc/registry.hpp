// serac_b200: helpers for registering named integrands (compiled-in kernel tables).
#pragma once

#include <new>

#include "common.hpp"
#include "serac_b200/kernel_table.cuh"
#include "serac_b200/physics_functors.hpp"

namespace sb {

using namespace serac;
using namespace serac::b200;

template <typename QF>
long store_blob(const QF& qf, void* out, size_t cap)
{
  if (cap < sizeof(QF)) return -1;
  std::memcpy(out, &qf, sizeof(QF));
  return long(sizeof(QF));
}

// signatures used by the physics modules:
//   SolidMechanics<p,dim>: H1<p,dim>( shape H1<1,dim>, u H1<p,dim>, a H1<p,dim> )     solid_mechanics.hpp:243
//   HeatTransfer<p,dim>  : H1<p>    ( shape H1<1,dim>, T H1<p>,     dT/dt H1<p> )     heat_transfer.hpp:209
template <Geometry g, int p>
struct SolidSig {
  static constexpr int dim = dimension_of(g);
  using test               = H1<p, dim>;
  using type               = Signature<g, p + 1, test, H1<1, dim>, H1<p, dim>, H1<p, dim>>;
  template <typename Material>
  using qf = ShapeAwareIntegrandWrapperWithState<MaterialStressFunctor<Material>, dim, test, H1<p, dim>, H1<p, dim>>;
};

template <Geometry g, int p>
struct ThermalSig {
  static constexpr int dim = dimension_of(g);
  using test               = H1<p>;
  using type               = Signature<g, p + 1, test, H1<1, dim>, H1<p>, H1<p>>;
  template <typename Material>
  using qf = ShapeAwareIntegrandWrapper<ThermalMaterialIntegrand<Material>, dim, test, H1<p>, H1<p>>;
};

}  // namespace sb
