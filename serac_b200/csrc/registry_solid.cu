// serac_b200: named solid-mechanics integrands -- what SolidMechanics::setMaterial registers with its Functional
// (src/serac/physics/solid_mechanics.hpp:879-894): ShapeAwareIntegrandWrapperWithState<MaterialStressFunctor<Material>>
// on H1<p,dim>(H1<1,dim> shape, H1<p,dim> u, H1<p,dim> a).
#include "registry.hpp"

namespace sb {
namespace {

using solid_mechanics::LinearIsotropic;
using solid_mechanics::NeoHookean;

template <typename S, typename Material>
long make_solid(const double* p, int n, void* out, size_t cap)
{
  if (n != 4) return -1;
  Material m{p[0], p[1], p[2]};
  auto     gn = (p[3] != 0.0) ? GeometricNonlinearities::On : GeometricNonlinearities::Off;
  typename S::template qf<Material> qf{MaterialStressFunctor<Material>(m, gn)};
  return store_blob(qf, out, cap);
}

#ifndef SB_REG_GEOM
#error "compile with -DSB_REG_GEOM=CUBE|SQUARE -DSB_REG_P=1|2|3 -DSB_REG_TAG=hex_p2"
#endif
#define SB_STR2(x) #x
#define SB_STR(x) SB_STR2(x)

using S = SolidSig<Geometry::SB_REG_GEOM, SB_REG_P>;

static Registrar r_neo({"solid_neohookean_" SB_STR(SB_REG_TAG),
                        make_kernel_table<S::type, S::qf<NeoHookean>, NeoHookean::State>(),
                        &make_solid<S, NeoHookean>, "{density, K, G, geometric_nonlinearity(0|1)}"});
static Registrar r_lin({"solid_linear_isotropic_" SB_STR(SB_REG_TAG),
                        make_kernel_table<S::type, S::qf<LinearIsotropic>, LinearIsotropic::State>(),
                        &make_solid<S, LinearIsotropic>, "{density, K, G, geometric_nonlinearity(0|1)}"});

#ifdef SB_REG_J2
// small-strain J2 plasticity with quadrature-point state (solid_material.hpp:201-259; State = plastic strain + eqps)
using J2Lin = solid_mechanics::J2SmallStrain<solid_mechanics::LinearHardening>;
template <typename SS>
long make_j2(const double* p, int n, void* out, size_t cap)
{
  if (n != 7) return -1;
  J2Lin m{p[0], p[1], solid_mechanics::LinearHardening{p[2], p[3]}, p[4], p[5]};
  auto  gn = (p[6] != 0.0) ? GeometricNonlinearities::On : GeometricNonlinearities::Off;
  typename SS::template qf<J2Lin> qf{MaterialStressFunctor<J2Lin>(m, gn)};
  return store_blob(qf, out, cap);
}
static Registrar r_j2({"solid_j2smallstrain_linear_" SB_STR(SB_REG_TAG),
                       make_kernel_table<S::type, S::qf<J2Lin>, J2Lin::State>(), &make_j2<S>,
                       "{E, nu, sigma_y, Hi, Hk, density, geometric_nonlinearity(0|1)}"});
#endif

#ifdef SB_REG_J2
// finite-deformation J2 (solid_material.hpp:261-330; State = inverse plastic distortion + eqps): BASELINE.json configs[2]
using J2FiniteLin = solid_mechanics::J2<solid_mechanics::LinearHardening>;
template <typename SS>
long make_j2_finite(const double* p, int n, void* out, size_t cap)
{
  if (n != 6) return -1;
  J2FiniteLin m{p[0], p[1], solid_mechanics::LinearHardening{p[2], p[3]}, p[4]};
  auto        gn = (p[5] != 0.0) ? GeometricNonlinearities::On : GeometricNonlinearities::Off;
  typename SS::template qf<J2FiniteLin> qf{MaterialStressFunctor<J2FiniteLin>(m, gn)};
  return store_blob(qf, out, cap);
}
static Registrar r_j2f({"solid_j2_linear_" SB_STR(SB_REG_TAG),
                        make_kernel_table<S::type, S::qf<J2FiniteLin>, J2FiniteLin::State>(), &make_j2_finite<S>,
                        "{E, nu, sigma_y, Hi, density, geometric_nonlinearity(0|1)}"});
#endif

#ifdef SB_REG_J2
// the other hardening laws (solid_material.hpp:154-198) and St. Venant-Kirchhoff (:61-92), H1 p <= 2 hexes
using J2Power = solid_mechanics::J2SmallStrain<solid_mechanics::PowerLawHardening>;
using J2Voce  = solid_mechanics::J2SmallStrain<solid_mechanics::VoceHardening>;
template <typename SS, typename Hardening>
long make_j2_hardening(const double* p, int n, void* out, size_t cap)
{
  if (n != 8) return -1;
  using M = solid_mechanics::J2SmallStrain<Hardening>;
  M    m{p[0], p[1], Hardening{p[2], p[3], p[4]}, p[5], p[6]};
  auto gn = (p[7] != 0.0) ? GeometricNonlinearities::On : GeometricNonlinearities::Off;
  typename SS::template qf<M> qf{MaterialStressFunctor<M>(m, gn)};
  return store_blob(qf, out, cap);
}
static Registrar r_j2p({"solid_j2smallstrain_powerlaw_" SB_STR(SB_REG_TAG),
                        make_kernel_table<S::type, S::qf<J2Power>, J2Power::State>(),
                        &make_j2_hardening<S, solid_mechanics::PowerLawHardening>,
                        "{E, nu, sigma_y, n, eps0, Hk, density, geometric_nonlinearity(0|1)}"});
static Registrar r_j2v({"solid_j2smallstrain_voce_" SB_STR(SB_REG_TAG),
                        make_kernel_table<S::type, S::qf<J2Voce>, J2Voce::State>(),
                        &make_j2_hardening<S, solid_mechanics::VoceHardening>,
                        "{E, nu, sigma_y, sigma_sat, strain_constant, Hk, density, geometric_nonlinearity(0|1)}"});
using solid_mechanics::StVenantKirchhoff;
static Registrar r_stvk({"solid_stvk_" SB_STR(SB_REG_TAG),
                         make_kernel_table<S::type, S::qf<StVenantKirchhoff>, StVenantKirchhoff::State>(),
                         &make_solid<S, StVenantKirchhoff>, "{density, K, G, geometric_nonlinearity(0|1)}"});
#endif

#ifdef SB_REG_AD
// the same integrand differentiated with dual numbers end to end (the reference's scheme; what an arbitrary user functor
// gets): kept to cross-check the tangent fast path
static Registrar r_neo_ad({"solid_neohookean_" SB_STR(SB_REG_TAG) "_ad",
                           make_kernel_table<S::type, S::qf<NeoHookean>, NeoHookean::State, 1, 0, false>(),
                           &make_solid<S, NeoHookean>, "{density, K, G, geometric_nonlinearity(0|1)}"});
#endif

}  // namespace
}  // namespace sb
