// serac_b200: tuning variants of the config-A integrand (not part of the default build: make TUNE=1).
// name: tune_e<EPB>_r<ROWS>_m<MINB>
#ifndef TUNE_SET
#define TUNE_SET 0
#endif
// sets 2 / 3: several elements per evaluation CTA, (element, point) flattened over the warps (216 threads = 7 warps at
// 96% lane use instead of one warp at 84%), at the price of block barriers
#if TUNE_SET == 2
#define SB_EVAL_CTA_E_HEX_Q3 8
#elif TUNE_SET == 3
#define SB_EVAL_CTA_E_HEX_Q3 16
#endif
#include "registry.hpp"

namespace sb {
namespace {
using solid_mechanics::NeoHookean;

template <int EPB>
struct TuneSig {
  static constexpr int dim = 3;
  using test               = H1<2, 3>;
  using type               = SignatureT<Geometry::CUBE, 3, EPB, test, H1<1, 3>, H1<2, 3>, H1<2, 3>>;
  using qf = ShapeAwareIntegrandWrapperWithState<MaterialStressFunctor<NeoHookean>, 3, test, H1<2, 3>, H1<2, 3>>;
};

template <int EPB>
long make_blob(const double* p, int n, void* out, size_t cap)
{
  if (n != 4) return -1;
  typename TuneSig<EPB>::qf qf{MaterialStressFunctor<NeoHookean>(
      NeoHookean{p[0], p[1], p[2]}, (p[3] != 0.0) ? GeometricNonlinearities::On : GeometricNonlinearities::Off)};
  return store_blob(qf, out, cap);
}

#define TUNE(E, R, M)                                                                                          \
  static Registrar tune_##E##_##R##_##M({"tune_e" #E "_r" #R "_m" #M,                                          \
                                         make_kernel_table<TuneSig<E>::type, TuneSig<E>::qf, Empty, R, M>(), \
                                         &make_blob<E>, "{density, K, G, gn}"});
#ifndef TUNE_SET
#define TUNE_SET 0
#endif
// EPB >= 1000: experiment switches in the thousands (SignatureT::tune): 1 = streaming stores of the derivatives,
// 2 = streaming loads of X_q / J_q
#if TUNE_SET == 0
TUNE(4, 1, 16)
TUNE(4, 1, 18)
#elif TUNE_SET == 1
TUNE(4, 1, 20)
#elif TUNE_SET == 2
TUNE(8, 1, 2)
#elif TUNE_SET == 3
TUNE(16, 1, 1)
#elif TUNE_SET == 10
// the streaming Jacobian action under symmetric tangent storage: E = (100 + resident CTAs asked of ptxas) * 1000 + elements
// per CTA (the evaluation kernel's variant is the default one throughout)
TUNE(103002, 1, 16)
TUNE(105002, 1, 16)
TUNE(106002, 1, 16)
TUNE(108002, 1, 16)
TUNE(102004, 1, 16)
TUNE(103004, 1, 16)
TUNE(104004, 1, 16)
TUNE(108001, 1, 16)
TUNE(112001, 1, 16)
TUNE(116001, 1, 16)
#endif
}  // namespace
}  // namespace sb
