// serac_b200: a named integrand with a DISCONTINUOUS parameter field -- `H1<p>(H1<p> u, L2<0> k)`, reaction-diffusion
// with a piecewise-constant coefficient: source (c0 + k) u, flux k^2 grad u.  This is how the physics modules take material
// parameters (Parameters<L2<0>>, src/serac/physics/solid_mechanics.hpp).  L2 spaces use the same nodal bases as H1
// (src/serac/numerics/functional/detail/hexahedron_L2.inl:90-130) and their element dofs are lexicographic already
// (element_restriction.cpp:262-268): the restriction table is the native dof list, unpermuted; order 0 has one node per
// element whose shape function is 1.
#include "registry.hpp"

namespace sb {
namespace {

#ifndef SB_REG_GEOM
#error "compile with -DSB_REG_GEOM=CUBE|TETRAHEDRON -DSB_REG_P=1|2 -DSB_REG_TAG=hex_p2"
#endif
#define SB_STR2(x) #x
#define SB_STR(x) SB_STR2(x)

constexpr Geometry g = Geometry::SB_REG_GEOM;
constexpr int      p = SB_REG_P;

struct ParameterizedDiffusion {
  double c0;
  template <typename X, typename U, typename K>
  SERAC_HOST_DEVICE auto operator()(double, X, U u, K k) const
  {
    auto uv = get<0>(u);
    auto du = get<1>(u);
    auto kv = get<0>(k);
    return serac::tuple{(c0 + kv) * uv, (kv * kv) * du};
  }
};

using Sig = Signature<g, p + 1, H1<p>, H1<p>, L2<0>>;

long make_param(const double* q, int n, void* out, size_t cap)
{
  if (n != 1) return -1;
  return store_blob(ParameterizedDiffusion{q[0]}, out, cap);
}

static Registrar r_param({"param_diffusion_" SB_STR(SB_REG_TAG), make_kernel_table<Sig, ParameterizedDiffusion, Nothing>(),
                          &make_param, "{c0}: source (c0 + k) u, flux k^2 grad u with k an L2<0> parameter field"});

}  // namespace
}  // namespace sb
