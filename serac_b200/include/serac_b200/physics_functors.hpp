// serac_b200: device-ready q-function functors of the physics modules that call Functional.
//
// These mirror, with the same names, members, argument meaning and maths, the functors the reference's callers build
// around a material model before handing it to Functional::AddDomainIntegral:
//   solid_mechanics::LinearIsotropic / NeoHookean / StVenantKirchhoff   src/serac/physics/materials/solid_material.hpp:24-127
//   solid_mechanics::J2SmallStrain / J2 + hardening laws                src/serac/physics/materials/solid_material.hpp:130-330
//   SolidMechanics::MaterialStressFunctor                               src/serac/physics/solid_mechanics.hpp:808-853
//   heat_transfer::LinearIsotropicConductor                             src/serac/physics/materials/thermal_material.hpp:20-65
//   HeatTransfer::ThermalMaterialIntegrand                              src/serac/physics/heat_transfer.hpp:381-405
//   detail::ShapeCorrection, ShapeAwareIntegrandWrapper[WithState]      src/serac/numerics/functional/shape_aware_functional.hpp:37-141,364-438
// Every operator() is SERAC_HOST_DEVICE and free of host-only constructs (no slic, no std::function, no host lambdas)
// so the evaluation kernels can inline them (SURVEY.md Appendix C lists what the reference still lacks).
#pragma once

#include "serac_b200/finite_element.hpp"
#include "serac_b200/tensor.hpp"
#include "serac_b200/scalar_solver.hpp"

namespace serac {

constexpr int SOURCE     = 0;
constexpr int FLUX       = 1;
constexpr int VALUE      = 0;
constexpr int DERIVATIVE = 1;

enum class GeometricNonlinearities
{
  On,
  Off
};

/// The two linear maps around a q-function at one quadrature point, folded together so that a tangent provider can apply
/// them in a single pass (an extension of the q-function concept; the reference differentiates *through* these maps
/// with dual numbers, src/serac/numerics/functional/domain_integral_kernels.hpp:181-213):
///   physical gradient = parent gradient . M      (M = inv(dX/dxi), times inv(I + grad p) under a shape-aware wrapper)
///   parent flux       = scale * physical flux . M^T,   parent source = scale * physical source
template <int dim>
struct ParentMap {
  tensor<double, dim, dim> M;
  double                   scale;
};

SB_HD double partial_of(double, int) { return 0.0; }
template <int N>
SB_HD double partial_of(const dual<N>& x, int k)
{
  return x.gradient[k];
}

/// does material M provide  stress_tangent_row(state, du_dX, sigma, b, Cb):  Cb[a][j][m] = d sigma[a][j] / d du_dX[b][m] ?
template <typename M, typename = void>
struct has_closed_form_tangent : std::false_type {
};
template <typename M>
struct has_closed_form_tangent<M, std::void_t<decltype(M::closed_form_tangent)>> : std::bool_constant<M::closed_form_tangent> {
};

/// Which parts of its arguments does a q-function read?  (extension; lets the evaluation kernel skip the interpolation of
/// values / gradients nobody looks at -- e.g. the gradient of the acceleration in SolidMechanics' integrand.)
/// A q-function may declare `static constexpr int argument_usage(int i)` (bit 0: value, bit 1: gradient of trial argument
/// i) and `static constexpr bool uses_position`; anything undeclared counts as used.
constexpr int USES_VALUE = 1, USES_GRADIENT = 2;
template <typename QF, typename = void>
struct qf_usage {
  static constexpr int  argument(int) { return USES_VALUE | USES_GRADIENT; }
  static constexpr bool position = true;
};
template <typename QF>
struct qf_usage<QF, std::void_t<decltype(QF::argument_usage(0))>> {
  static constexpr int  argument(int i) { return QF::argument_usage(i); }
  static constexpr bool position = QF::uses_position;
};

/// does q-function QF provide flux_tangent(integral_constant<WRT>, ...) for its argument WRT ?
template <typename QF, int WRT, typename = void>
struct provides_flux_tangent : std::false_type {
};
template <typename QF, int WRT>
struct provides_flux_tangent<QF, WRT, std::void_t<decltype(QF::template has_flux_tangent<WRT>)>>
    : std::bool_constant<QF::template has_flux_tangent<WRT>> {
};

/// does q-function QF vouch that the flux tangent it provides for argument WRT has the major symmetry
/// d flux[a][k] / d grad[b][l] = d flux[b][l] / d grad[a][k] ?  Then the kernels store the upper triangle only
/// (DerivativeLayout::sym, domain_integral.cuh): 45 instead of 81 doubles per quadrature point in 3-D.
template <typename QF, int WRT, typename = void>
struct provides_symmetric_flux_tangent : std::false_type {
};
#ifndef SB_SYMMETRIC_TANGENT
#define SB_SYMMETRIC_TANGENT 1  // 0: every tangent is stored in full (for A/B measurements)
#endif
template <typename QF, int WRT>
struct provides_symmetric_flux_tangent<QF, WRT, std::void_t<decltype(QF::template has_symmetric_flux_tangent<WRT>)>>
    : std::bool_constant<(SB_SYMMETRIC_TANGENT != 0) && QF::template has_symmetric_flux_tangent<WRT>> {
};

namespace solid_mechanics {

/// linear isotropic elasticity (solid_material.hpp:24-50)
struct LinearIsotropic {
  using State = Empty;
  template <typename T, int dim>
  SERAC_HOST_DEVICE auto operator()(State&, const tensor<T, dim, dim>& du_dX) const
  {
    auto I       = Identity<dim>();
    auto lambda  = K - (2.0 / 3.0) * G;
    auto epsilon = 0.5 * (transpose(du_dX) + du_dX);
    return lambda * tr(epsilon) * I + 2.0 * G * epsilon;
  }
  /// closed-form d sigma / d du_dX, row b: lambda d_aj d_bm + G (d_ab d_jm + d_am d_jb)
  static constexpr bool closed_form_tangent = true;
  /// ... which is the same under (a, j) <-> (b, m): without geometric nonlinearity the flux tangent is symmetric
  static constexpr bool symmetric_small_strain_tangent = true;
  template <int dim>
  SERAC_HOST_DEVICE void stress_tangent_row(const State&, const tensor<double, dim, dim>&, const tensor<double, dim, dim>&,
                                            int b, tensor<double, dim, dim, dim>& Cb) const
  {
    const double lambda = K - (2.0 / 3.0) * G;
    SB_UNROLL
    for (int a = 0; a < dim; a++) {
      SB_UNROLL
      for (int j = 0; j < dim; j++) {
        SB_UNROLL
        for (int m = 0; m < dim; m++) {
          Cb[a][j][m] = ((a == j && b == m) ? lambda : 0.0) + ((a == b && j == m) ? G : 0.0) + ((a == m && j == b) ? G : 0.0);
        }
      }
    }
  }
  double density;
  double K;
  double G;
};

template <typename T, int dim>
SERAC_HOST_DEVICE auto greenStrain(const tensor<T, dim, dim>& grad_u)
{
  return 0.5 * (grad_u + transpose(grad_u) + dot(transpose(grad_u), grad_u));
}

/// St. Venant-Kirchhoff (solid_material.hpp:61-92)
struct StVenantKirchhoff {
  using State = Empty;
  template <typename T, int dim>
  SERAC_HOST_DEVICE auto operator()(State&, const tensor<T, dim, dim>& grad_u) const
  {
    auto       I     = Identity<dim>();
    auto       F     = grad_u + I;
    const auto E     = greenStrain(grad_u);
    const auto S     = K * tr(E) * I + 2.0 * G * dev(E);
    const auto P     = dot(F, S);
    const auto sigma = dot(P, transpose(F)) / det(F);
    return sigma;
  }
  double density;
  double K;
  double G;
};

/// compressible neo-Hookean (solid_material.hpp:98-127)
struct NeoHookean {
  using State = Empty;
  template <typename T, int dim>
  SERAC_HOST_DEVICE auto operator()(State&, const tensor<T, dim, dim>& du_dX) const
  {
    auto I         = Identity<dim>();
    auto lambda    = K - (2.0 / 3.0) * G;
    auto B_minus_I = du_dX * transpose(du_dX) + transpose(du_dX) + du_dX;
    auto J_minus_1 = detApIm1(du_dX);
    auto J         = J_minus_1 + 1;
    return (lambda * log1p(J_minus_1) * I + G * B_minus_I) / J;
  }
  /// first Piola stress P = lambda log J F^-T + G (F - F^-T):
  ///   dP[a][i]/dF[b][m] = lambda Finv[i][a] Finv[m][b] + (G - lambda log J) Finv[i][b] Finv[m][a] + G d_ab d_im
  static constexpr bool closed_form_piola_tangent = true;
  template <int dim>
  SERAC_HOST_DEVICE tensor<double, 3> piola_tangent_coefficients(const State&, const tensor<double, dim, dim>& du_dX) const
  {
    const double lambda = K - (2.0 / 3.0) * G;
    return tensor<double, 3>{{lambda, G - lambda * log1p(detApIm1(du_dX)), G}};
  }
  /// closed-form d sigma / d du_dX, row b.  With F = I + H, B = F F^T, J = det F:
  ///   d sigma[a][j] / d H[b][m] = ( lambda d_aj Finv[m][b] + G (d_ab F[j][m] + F[a][m] d_jb) ) / J - sigma[a][j] Finv[m][b]
  static constexpr bool closed_form_tangent = true;
  template <int dim>
  SERAC_HOST_DEVICE void stress_tangent_row(const State&, const tensor<double, dim, dim>& du_dX,
                                            const tensor<double, dim, dim>& sigma, int b,
                                            tensor<double, dim, dim, dim>& Cb) const
  {
    const double lambda = K - (2.0 / 3.0) * G;
    const auto   F      = du_dX + Identity<dim>();
    const auto   Finv   = inv(F);
    const double rJ     = 1.0 / (detApIm1(du_dX) + 1);
    SB_UNROLL
    for (int a = 0; a < dim; a++) {
      SB_UNROLL
      for (int j = 0; j < dim; j++) {
        SB_UNROLL
        for (int m = 0; m < dim; m++) {
          double v = -sigma[a][j] * Finv[m][b];
          if (a == j) v += lambda * rJ * Finv[m][b];
          if (a == b) v += G * rJ * F[j][m];
          if (j == b) v += G * rJ * F[a][m];
          Cb[a][j][m] = v;
        }
      }
    }
  }
  double density;
  double K;
  double G;
};

/// sigma_y + Hi * eqps (solid_material.hpp:132-149)
struct LinearHardening {
  double sigma_y;
  double Hi;
  template <typename T>
  SERAC_HOST_DEVICE auto operator()(const T accumulated_plastic_strain) const
  {
    return sigma_y + Hi * accumulated_plastic_strain;
  }
};

/// sigma_y (1 + eqps/eps0)^(1/n) (solid_material.hpp:154-173)
struct PowerLawHardening {
  double sigma_y;
  double n;
  double eps0;
  template <typename T>
  SERAC_HOST_DEVICE auto operator()(const T accumulated_plastic_strain) const
  {
    return sigma_y * pow(1.0 + accumulated_plastic_strain / eps0, 1.0 / n);
  }
};

/// Voce saturation hardening (solid_material.hpp:180-198)
struct VoceHardening {
  double sigma_y;
  double sigma_sat;
  double strain_constant;
  template <typename T>
  SERAC_HOST_DEVICE auto operator()(const T accumulated_plastic_strain) const
  {
    return sigma_sat - (sigma_sat - sigma_y) * exp(-1.0 * accumulated_plastic_strain / strain_constant);
  }
};

/// small-strain J2 plasticity with isotropic + linear kinematic hardening (solid_material.hpp:201-259)
template <typename HardeningType>
struct J2SmallStrain {
  static constexpr int    dim = 3;
  static constexpr double tol = 1e-10;

  double        E;
  double        nu;
  HardeningType hardening;
  double        Hk;
  double        density;

  struct State {
    tensor<double, dim, dim> plastic_strain;
    double                   accumulated_plastic_strain;
  };

  /// return-map residual: trial_q - (3G + Hk) d - flow_stress(eqps_old + d)
  struct Residual {
    double        eqps_old, G, Hk;
    HardeningType hardening;
    template <typename T1, typename T2>
    SERAC_HOST_DEVICE auto operator()(const T1& delta_eqps, const T2& trial_q) const
    {
      return trial_q - (3.0 * G + Hk) * delta_eqps - hardening(eqps_old + delta_eqps);
    }
  };

  template <typename T>
  SERAC_HOST_DEVICE auto operator()(State& state, const tensor<T, dim, dim>& du_dX) const
  {
    auto         I = Identity<dim>();
    const double K = E / (3.0 * (1.0 - 2.0 * nu));
    const double G = 0.5 * E / (1.0 + nu);

    // (i) elastic predictor
    auto el_strain = sym(du_dX) - state.plastic_strain;
    auto p         = K * tr(el_strain);
    auto s         = 2.0 * G * dev(el_strain);
    auto sigma_b   = (2.0 / 3.0 * Hk) * state.plastic_strain;
    auto eta       = s - sigma_b;
    auto q         = ::sqrt(1.5) * norm(eta);

    // (ii) admissibility
    const double eqps_old = state.accumulated_plastic_strain;
    Residual     residual{eqps_old, G, Hk, hardening};
    if (residual(0.0, get_value(q)) > tol * hardening.sigma_y) {
      // (iii) return mapping
      ScalarSolverOptions opts{0, tol * hardening.sigma_y, 25};
      double              lower_bound = 0.0;
      double              upper_bound = (get_value(q) - hardening(eqps_old)) / (3.0 * G + Hk);
      auto                delta_eqps  = get<0>(solve_scalar_equation(residual, 0.0, lower_bound, upper_bound, opts, q));

      auto Np = 1.5 * eta / q;

      s = s - 2.0 * G * delta_eqps * Np;
      state.accumulated_plastic_strain += get_value(delta_eqps);
      state.plastic_strain += get_value(delta_eqps) * get_value(Np);
    }
    return s + p * I;
  }
};

/// finite-deformation J2 plasticity, logarithmic strain, isotropic hardening (solid_material.hpp:261-330):
/// multiplicative split F = Fe Fp with the inverse plastic distortion and the equivalent plastic strain as state
template <typename HardeningType>
struct J2 {
  static constexpr int    dim = 3;
  static constexpr double tol = 1e-10;

  double        E;
  double        nu;
  HardeningType hardening;
  double        density;

  struct State {
    /// inverse of the plastic distortion; State{} is the virgin material, as in the reference (solid_material.hpp:273)
    tensor<double, dim, dim> Fpinv = DenseIdentity<dim>();
    double                   accumulated_plastic_strain;  ///< uniaxial equivalent plastic strain
  };

  struct Residual {
    double        eqps_old, G;
    HardeningType hardening;
    template <typename T1, typename T2>
    SERAC_HOST_DEVICE auto operator()(const T1& delta_eqps, const T2& trial_mises) const
    {
      return trial_mises - 3.0 * G * delta_eqps - hardening(eqps_old + delta_eqps);
    }
  };

  template <typename T>
  SERAC_HOST_DEVICE auto operator()(State& state, const tensor<T, dim, dim>& du_dX) const
  {
    auto         I = Identity<dim>();
    const double K = E / (3.0 * (1.0 - 2.0 * nu));
    const double G = 0.5 * E / (1.0 + nu);

    // (i) elastic predictor in logarithmic strain
    auto F  = du_dX + I;
    auto Fe = dot(F, state.Fpinv);
    auto Ee = 0.5 * log_symm(dot(transpose(Fe), Fe));
    auto p  = K * tr(Ee);
    auto s  = 2.0 * G * dev(Ee);
    auto q  = ::sqrt(1.5) * norm(s);

    // (ii) admissibility
    const double eqps_old = state.accumulated_plastic_strain;
    Residual     residual{eqps_old, G, hardening};
    if (residual(0.0, get_value(q)) > tol * hardening.sigma_y) {
      // (iii) return mapping: radial in the Mandel stress, exponential update of the plastic distortion
      ScalarSolverOptions opts{0, tol * hardening.sigma_y, 25};
      double              lower_bound = 0.0;
      double              upper_bound = (get_value(q) - hardening(eqps_old)) / (3.0 * G);
      auto                delta_eqps  = get<0>(solve_scalar_equation(residual, 0.0, lower_bound, upper_bound, opts, q));

      auto Np = 1.5 * s / q;

      s      = s - 2.0 * G * delta_eqps * Np;
      auto A = exp_symm(-1.0 * delta_eqps * Np);
      Fe     = dot(Fe, A);
      state.accumulated_plastic_strain += get_value(delta_eqps);
      state.Fpinv = dot(state.Fpinv, get_value(A));
    }
    auto M   = s + p * I;  // Mandel stress -> Cauchy
    auto FeT = transpose(Fe);
    return dot(dot(inv(FeT), M), FeT) / det(F);
  }
};

}  // namespace solid_mechanics

namespace heat_transfer {

/// constant isotropic conductor (thermal_material.hpp:20-65); returns {volumetric heat capacity, heat flux}
struct LinearIsotropicConductor {
  LinearIsotropicConductor(double density = 1.0, double specific_heat_capacity = 1.0, double conductivity = 1.0)
      : density_(density), specific_heat_capacity_(specific_heat_capacity), conductivity_(conductivity)
  {
  }
  template <typename T1, typename T2, typename T3>
  SERAC_HOST_DEVICE auto operator()(const T1&, const T2&, const T3& temperature_gradient) const
  {
    return serac::tuple{density_ * specific_heat_capacity_, -1.0 * conductivity_ * temperature_gradient};
  }
  double density_;
  double specific_heat_capacity_;
  double conductivity_;
};

}  // namespace heat_transfer

// ------------------------------------------------------------------------------------------------
// thermomechanics (BASELINE.json configs[3]): one material evaluated by both physics modules
// ------------------------------------------------------------------------------------------------
/// Green-Saint Venant thermoelastic model (materials/green_saint_venant_thermoelastic.hpp:26-92): returns
/// tuple{Cauchy stress, volumetric heat capacity, internal heat source, heat flux}
struct GreenSaintVenantThermoelasticMaterial {
  double density;    ///< density
  double E;          ///< Young's modulus
  double nu;         ///< Poisson's ratio
  double C_v;        ///< volumetric heat capacity
  double alpha;      ///< thermal expansion coefficient
  double theta_ref;  ///< datum temperature for thermal expansion
  double k;          ///< thermal conductivity

  struct State {
    double strain_trace;  ///< trace of the Green-Saint Venant strain tensor
  };

  template <typename T1, typename T2, typename T3>
  SERAC_HOST_DEVICE auto operator()(State& state, const tensor<T1, 3, 3>& grad_u, T2 theta,
                                    const tensor<T3, 3>& grad_theta) const
  {
    const double K    = E / (3.0 * (1.0 - 2.0 * nu));
    const double G    = 0.5 * E / (1.0 + nu);
    const auto   I    = Identity<3>();
    auto         F    = grad_u + I;
    const auto   Eg   = solid_mechanics::greenStrain(grad_u);
    const auto   trEg = tr(Eg);

    const auto S     = 2.0 * G * dev(Eg) + K * (trEg - 3.0 * alpha * (theta - theta_ref)) * I;
    const auto P     = dot(F, S);
    const auto sigma = dot(P, transpose(F)) / det(F);

    const auto s0 = -3 * K * alpha * theta * (trEg - state.strain_trace);  // internal heat source
    const auto q0 = -k * grad_theta;                                       // heat flux

    state.strain_trace = get_value(trEg);
    return serac::tuple{sigma, C_v, s0, q0};
  }
};

namespace thermomechanics {

/// makes a thermomechanical material usable by HeatTransfer (thermomechanics.hpp:272-310); the State is a throw-away
/// local, as in the reference
template <typename ThermalMechanicalMaterial>
struct ThermalMaterialInterface {
  using State = typename ThermalMechanicalMaterial::State;
  ThermalMechanicalMaterial mat;
  SERAC_HOST_DEVICE ThermalMaterialInterface(const ThermalMechanicalMaterial& m) : mat(m) {}

  template <typename T1, typename T2, typename T3, typename T4, typename... param_types>
  SERAC_HOST_DEVICE auto operator()(const T1& /* x */, const T2& temperature, const T3& temperature_gradient,
                                    const T4& displacement, param_types... parameters) const
  {
    State state{};
    auto  du_dX = get<DERIVATIVE>(displacement);
    auto  resp  = mat(state, du_dX, temperature, temperature_gradient, parameters...);
    return serac::tuple{get<1>(resp), get<3>(resp)};
  }
};

/// makes a thermomechanical material usable by SolidMechanics (thermomechanics.hpp:318-352)
template <typename ThermalMechanicalMaterial>
struct MechanicalMaterialInterface {
  using State = typename ThermalMechanicalMaterial::State;
  static constexpr bool     parameterized = true;
  ThermalMechanicalMaterial mat;
  double                    density;
  SERAC_HOST_DEVICE MechanicalMaterialInterface(ThermalMechanicalMaterial m) : mat(m), density(m.density) {}

  template <typename T1, typename T2, typename... param_types>
  SERAC_HOST_DEVICE auto operator()(State& state, const T1& displacement_gradient, const T2& temperature,
                                    param_types... parameters) const
  {
    auto theta     = get<VALUE>(temperature);
    auto dtheta_dX = get<DERIVATIVE>(temperature);
    auto resp      = mat(state, displacement_gradient, theta, dtheta_dX, parameters...);
    return get<0>(resp);
  }
};

}  // namespace thermomechanics

/// does material M declare that its small-strain moduli d sigma[a][j] / d du_dX[b][m] have the major symmetry ?
template <typename M, typename = void>
struct has_symmetric_small_strain_tangent : std::false_type {
};
template <typename M>
struct has_symmetric_small_strain_tangent<M, std::void_t<decltype(M::symmetric_small_strain_tangent)>>
    : std::bool_constant<M::symmetric_small_strain_tangent> {
};

/// does material M provide piola_tangent_coefficients(state, du_dX) -> {c1, c2, c3} with
///   d P[a][i] / d F[b][m] = c1 Finv[i][a] Finv[m][b] + c2 Finv[i][b] Finv[m][a] + c3 d_ab d_im      (P = J sigma F^-T) ?
template <typename M, typename = void>
struct has_piola_tangent_coefficients : std::false_type {
};
template <typename M>
struct has_piola_tangent_coefficients<M, std::void_t<decltype(M::closed_form_piola_tangent)>>
    : std::bool_constant<M::closed_form_piola_tangent> {
};

/// materials that take parameter fields after du_dX (e.g. MechanicalMaterialInterface: the temperature) are
/// differentiated with dual numbers end to end: the tangent fast path below is written for (displacement, acceleration)
template <typename M, typename = void>
struct is_parameterized_material : std::false_type {
};
template <typename M>
struct is_parameterized_material<M, std::void_t<decltype(M::parameterized)>> : std::bool_constant<M::parameterized> {
};

/// how MaterialStressFunctor knows its geometric-nonlinearity flag: at run time (the reference's member), or fixed at
/// compile time (the kernel launcher converts a Dynamic functor into the matching static variant before the launch, so
/// the device code carries no dead branch and its register allocation is that of the live one)
enum class FlagMode
{
  Dynamic,
  StaticOff,
  StaticOn
};

/// q-function SolidMechanics::setMaterial registers (solid_mechanics.hpp:808-853)
template <typename Material, FlagMode MODE = FlagMode::Dynamic>
struct MaterialStressFunctor {
  SERAC_HOST_DEVICE MaterialStressFunctor(Material material, GeometricNonlinearities gn)
      : material_(material), geom_nonlin_(gn)
  {
  }
  Material                material_;
  GeometricNonlinearities geom_nonlin_;

  // --- run-time flag -> compile-time variant (see FlagMode) ---
  static constexpr int num_variants = (MODE == FlagMode::Dynamic) ? 2 : 1;
  template <int V>
  using variant = MaterialStressFunctor<Material, V == 0 ? FlagMode::StaticOff : FlagMode::StaticOn>;
  int variant_index() const { return geom_nonlin_ == GeometricNonlinearities::On ? 1 : 0; }
  template <int V>
  variant<V> as_variant() const
  {
    return variant<V>(material_, geom_nonlin_);
  }

  /// operator() reads the displacement gradient, the acceleration value and nothing of the position
  static constexpr int  argument_usage(int i) { return i == 0 ? USES_GRADIENT : i == 1 ? USES_VALUE : (USES_VALUE | USES_GRADIENT); }
  static constexpr bool uses_position = false;

  SERAC_HOST_DEVICE bool nonlinear() const
  {
    if constexpr (MODE == FlagMode::StaticOn) return true;
    if constexpr (MODE == FlagMode::StaticOff) return false;
    return geom_nonlin_ == GeometricNonlinearities::On;
  }

  template <typename X, typename State, typename Displacement, typename Acceleration, typename... Params>
  SERAC_HOST_DEVICE auto operator()(double, X, State& state, Displacement displacement, Acceleration acceleration,
                                    Params... params) const
  {
    auto du_dX   = get<DERIVATIVE>(displacement);
    auto d2u_dt2 = get<VALUE>(acceleration);

    auto stress = material_(state, du_dX, params...);

    constexpr int dim = std::decay_t<decltype(du_dX)>::first_dim;
    auto          dx_dX = 0.0 * du_dX + Identity<dim>();
    if (nonlinear()) {
      dx_dX = dx_dX + du_dX;
    }
    auto flux = dot(stress, transpose(inv(dx_dX))) * det(dx_dX);
    return serac::tuple{material_.density * d2u_dt2, flux};
  }

  /// Tangent fast path (extension): the derivative w.r.t. the displacement (argument 0) touches only the flux and only
  /// through du_dX, so instead of pushing dual numbers through the Cauchy -> Piola transform and the surrounding linear
  /// maps, differentiate the material alone (closed form when the material has one, else forward-mode AD one gradient
  /// row at a time) and apply everything else analytically.  With F = I + H, P = J sigma F^-T, A = M F^-1, Pp = P M^T:
  ///   parent flux                               = scale Pp
  ///   d(parent flux)[a][k]/d(parent grad)[b][l] = scale ( J sum_jm A[k][j] M[l][m] C[a][j][b][m] - A[k][b] Pp[a][l] + Pp[a][k] A[l][b] )
  /// (geometric nonlinearity off: P = sigma, A = M and the last two terms vanish).  Materials that know their first
  /// Piola tangent as c1 F^-T (x) F^-T + c2 F^-T [x] F^-T + c3 I [x] I shortcut further:
  ///   d(parent flux)[a][k]/d(parent grad)[b][l] = scale ( c1 A[k][a] A[l][b] + c2 A[k][b] A[l][a] + c3 d_ab (M M^T)[k][l] )
  /// Same mathematics as the reference's dual-number evaluation of operator() above, to rounding.
  template <int WRT>
  static constexpr bool has_flux_tangent = (WRT == 0) && !is_parameterized_material<Material>::value;
  /// with geometric nonlinearity on, a material that gives its first Piola tangent in the three-coefficient form above
  /// is hyperelastic and the parent tangent c1 A[k][a] A[l][b] + c2 A[k][b] A[l][a] + c3 d_ab (M M^T)[k][l] is
  /// symmetric under (a, k) <-> (b, l) term by term; known at compile time only for the static variant of the flag
  /// (geometric nonlinearity off: flux = sigma, and the parent tangent scale sum_jm M[k][j] M[l][m] C[a][j][b][m] is
  /// symmetric whenever the small-strain moduli C are)
  template <int WRT>
  static constexpr bool has_symmetric_flux_tangent =
      has_flux_tangent<WRT> &&
      ((MODE == FlagMode::StaticOn && has_piola_tangent_coefficients<Material>::value) ||
       (MODE == FlagMode::StaticOff && has_symmetric_small_strain_tangent<Material>::value));

  template <typename X, typename State, int dim, typename Sink, typename Displacement, typename Acceleration>
  SERAC_HOST_DEVICE void flux_tangent(std::integral_constant<int, 0>, double, const X&, State& state,
                                      const ParentMap<dim>& map, tensor<double, dim>& source,
                                      tensor<double, dim, dim>& flux, Sink&& sink, const Displacement& displacement,
                                      const Acceleration& acceleration) const
  {
    const tensor<double, dim, dim> H = get<DERIVATIVE>(displacement);
    source                           = (material_.density * map.scale) * get<VALUE>(acceleration);
    const tensor<double, dim, dim>& M = map.M;

    if constexpr (has_piola_tangent_coefficients<Material>::value) {
      if (nonlinear()) {
        const auto   c     = material_.piola_tangent_coefficients(state, H);
        const auto   sigma = material_(state, H);
        const auto   F     = H + Identity<dim>();
        const auto   A     = dot(M, inv(F));
        flux               = dot(sigma, transpose(A)) * (det(F) * map.scale);
        const double c1 = c[0] * map.scale, c2 = c[1] * map.scale, c3 = c[2] * map.scale;
        tensor<double, dim, dim> Q{};  // c3 M M^T (symmetric)
        SB_UNROLL
        for (int k = 0; k < dim; k++) {
          SB_UNROLL
          for (int l = k; l < dim; l++) {
            double t = M[k][0] * M[l][0];
            SB_UNROLL
            for (int m = 1; m < dim; m++) t += M[k][m] * M[l][m];
            Q[k][l] = Q[l][k] = c3 * t;
          }
        }
        SB_UNROLL
        for (int a = 0; a < dim; a++) {
          SB_UNROLL
          for (int k = 0; k < dim; k++) {
            const double c1Aka = c1 * A[k][a];
            SB_UNROLL
            for (int b = 0; b < dim; b++) {
              const double c2Akb = c2 * A[k][b];
              SB_UNROLL
              for (int l = 0; l < dim; l++) {
                double d = c1Aka * A[l][b] + c2Akb * A[l][a];
                if (a == b) d += Q[k][l];
                sink(a, k, b, l, d);
              }
            }
          }
        }
        return;
      }
    }

    const bool  nonlin    = nonlinear();
    const State state_old = state;
    tensor<double, dim, dim> sigma{};
    if constexpr (has_closed_form_tangent<Material>::value) {
      sigma = material_(state, H);
    }

    // A = M Finv, As = scale J A, sPp = scale Pp = sigma As^T
    tensor<double, dim, dim> A = M, As = map.scale * M, sPp{};
    if (nonlin) {
      const auto F = H + Identity<dim>();
      A            = dot(M, inv(F));
      As           = (det(F) * map.scale) * A;
    }
    SB_UNROLL
    for (int b = 0; b < dim; b++) {
      // Cb[a][j][m] = d sigma[a][j] / d H[b][m]
      tensor<double, dim, dim, dim> Cb{};
      if constexpr (has_closed_form_tangent<Material>::value) {
        material_.stress_tangent_row(state_old, H, sigma, b, Cb);
      } else {
        tensor<dual<dim>, dim, dim> Hd{};
        SB_UNROLL
        for (int i = 0; i < dim; i++) {
          SB_UNROLL
          for (int m = 0; m < dim; m++) {
            Hd[i][m] = make_dual_constant<dim>(H[i][m]);
            if (i == b) Hd[i][m].gradient[m] = 1.0;
          }
        }
        State st = state_old;
        auto  sd = material_(st, Hd);
        SB_UNROLL
        for (int a = 0; a < dim; a++) {
          SB_UNROLL
          for (int j = 0; j < dim; j++) {
            if (b == 0) sigma[a][j] = get_value(sd[a][j]);
            SB_UNROLL
            for (int m = 0; m < dim; m++) Cb[a][j][m] = partial_of(sd[a][j], m);
          }
        }
        if (b == 0) state = st;
      }
      if (b == 0) {
        sPp  = dot(sigma, transpose(As));  // (the AD path only has sigma after its first pass)
        flux = sPp;
      }
      SB_UNROLL
      for (int a = 0; a < dim; a++) {
        // T[j][l] = sum_m Cb[a][j][m] M[l][m]
        tensor<double, dim, dim> T{};
        SB_UNROLL
        for (int j = 0; j < dim; j++) {
          SB_UNROLL
          for (int l = 0; l < dim; l++) {
            double t = Cb[a][j][0] * M[l][0];
            SB_UNROLL
            for (int m = 1; m < dim; m++) t += Cb[a][j][m] * M[l][m];
            T[j][l] = t;
          }
        }
        SB_UNROLL
        for (int k = 0; k < dim; k++) {
          SB_UNROLL
          for (int l = 0; l < dim; l++) {
            double d = As[k][0] * T[0][l];
            SB_UNROLL
            for (int j = 1; j < dim; j++) d += As[k][j] * T[j][l];
            if (nonlin) d += sPp[a][k] * A[l][b] - A[k][b] * sPp[a][l];
            sink(a, k, b, l, d);
          }
        }
      }
    }
  }
};

/// q-function HeatTransfer::setMaterial registers (heat_transfer.hpp:381-405)
template <typename MaterialType>
struct ThermalMaterialIntegrand {
  SERAC_HOST_DEVICE ThermalMaterialIntegrand(MaterialType material) : material_(material) {}
  template <typename X, typename T, typename dT_dt, typename... Params>
  SERAC_HOST_DEVICE auto operator()(double, X x, T temperature, dT_dt dtemp_dt, Params... params) const
  {
    auto u     = get<VALUE>(temperature);
    auto du_dX = get<DERIVATIVE>(temperature);
    auto du_dt = get<VALUE>(dtemp_dt);
    auto resp  = material_(x, u, du_dX, params...);
    return serac::tuple{get<0>(resp) * du_dt, -1.0 * get<1>(resp)};
  }
  MaterialType material_;
};

namespace detail {

/// pull-backs for a shape displacement field p: x = X + p (shape_aware_functional.hpp:37-141)
template <int dim, typename shape_type>
struct ShapeCorrection {
  using jacobian_type = std::decay_t<decltype(get<DERIVATIVE>(std::declval<shape_type>()))>;
  using scalar_type   = typename jacobian_type::value_type;

  SERAC_HOST_DEVICE ShapeCorrection(Dimension<dim>, const shape_type& p)
      : J_(Identity<dim>() + get<DERIVATIVE>(p)), detJ_(det(J_)), inv_J_(inv(J_)), inv_JT_(transpose(inv_J_))
  {
  }

  /// du/dx' = du/dX (dx'/dX)^-1 for H1 / L2 trial arguments
  template <typename space_type, typename trial_type>
  SERAC_HOST_DEVICE auto modify_trial_argument(space_type, const trial_type& u) const
  {
    static_assert(space_type::family == Family::H1 || space_type::family == Family::L2,
                  "shape-aware pull-back implemented for H1 and L2 trial spaces");
    auto trial_derivative = dot(get<DERIVATIVE>(u), inv_J_);
    return serac::tuple{get<VALUE>(u), trial_derivative};
  }

  /// map the q-function's (source, flux) back to the reference configuration
  template <typename test_space, typename q_func_type>
  SERAC_HOST_DEVICE auto modify_shape_aware_qf_return(test_space, const q_func_type& v) const
  {
    if constexpr (std::is_same_v<test_space, double>) {
      return detJ_ * v;
    } else {
      auto modified_flux   = dot(get<FLUX>(v), inv_JT_) * detJ_;
      auto modified_source = get<SOURCE>(v) * detJ_;
      return serac::tuple{modified_source, modified_flux};
    }
  }

  tensor<scalar_type, dim, dim> J_;
  scalar_type                   detJ_;
  tensor<scalar_type, dim, dim> inv_J_;
  tensor<scalar_type, dim, dim> inv_JT_;
};

}  // namespace detail

/// what ShapeAwareFunctional::AddDomainIntegral wraps every integrand in (shape_aware_functional.hpp:364-438,456-468):
/// argument 0 is the shape displacement, the remaining arguments are the user's trial functions.
template <typename Integrand, int dim, typename test_space, typename... trial_spaces>
struct ShapeAwareIntegrandWrapper {
  SERAC_HOST_DEVICE ShapeAwareIntegrandWrapper(Integrand integrand) : integrand_(integrand) {}
  Integrand integrand_;
  /// the shape displacement's gradient is always read, its value only moves the position handed to the integrand
  static constexpr int argument_usage(int i)
  {
    return i == 0 ? (USES_GRADIENT | (qf_usage<Integrand>::position ? USES_VALUE : 0)) : qf_usage<Integrand>::argument(i - 1);
  }
  static constexpr bool uses_position = qf_usage<Integrand>::position;

  template <typename PositionType, typename ShapeValueType, typename... QFuncArgs>
  SERAC_HOST_DEVICE auto operator()(double time, PositionType x, ShapeValueType shape_val,
                                    QFuncArgs... qfunc_args) const
  {
    detail::ShapeCorrection<dim, ShapeValueType> shape_correction(Dimension<dim>{}, shape_val);
    auto x_mod = serac::tuple{get<VALUE>(x) + get<VALUE>(shape_val),
                              get<DERIVATIVE>(x) + get<DERIVATIVE>(shape_val) * get<DERIVATIVE>(x)};
    auto unmodified =
        integrand_(time, x_mod, shape_correction.modify_trial_argument(trial_spaces{}, qfunc_args)...);
    return shape_correction.modify_shape_aware_qf_return(test_space{}, unmodified);
  }
};

/// same, for integrands with quadrature-point state
template <typename Integrand, int dim, typename test_space, typename... trial_spaces>
struct ShapeAwareIntegrandWrapperWithState {
  SERAC_HOST_DEVICE ShapeAwareIntegrandWrapperWithState(Integrand integrand) : integrand_(integrand) {}
  Integrand integrand_;
  static constexpr int argument_usage(int i)
  {
    return i == 0 ? (USES_GRADIENT | (qf_usage<Integrand>::position ? USES_VALUE : 0)) : qf_usage<Integrand>::argument(i - 1);
  }
  static constexpr bool uses_position = qf_usage<Integrand>::position;

  template <typename PositionType, typename StateType, typename ShapeValueType, typename... QFuncArgs>
  SERAC_HOST_DEVICE auto operator()(double time, PositionType x, StateType& state, ShapeValueType shape_val,
                                    QFuncArgs... qfunc_args) const
  {
    detail::ShapeCorrection<dim, ShapeValueType> shape_correction(Dimension<dim>{}, shape_val);
    auto x_mod = serac::tuple{get<VALUE>(x) + get<VALUE>(shape_val),
                              get<DERIVATIVE>(x) + get<DERIVATIVE>(shape_val) * get<DERIVATIVE>(x)};
    auto unmodified =
        integrand_(time, x_mod, state, shape_correction.modify_trial_argument(trial_spaces{}, qfunc_args)...);
    return shape_correction.modify_shape_aware_qf_return(test_space{}, unmodified);
  }

  // run-time flag -> compile-time variant forwarding (see FlagMode)
  template <typename I, typename = void>
  struct variants_of {
    static constexpr int n = 1;
  };
  template <typename I>
  struct variants_of<I, std::void_t<decltype(I::num_variants)>> {
    static constexpr int n = I::num_variants;
  };
  static constexpr int num_variants = variants_of<Integrand>::n;
  template <int V>
  using variant =
      ShapeAwareIntegrandWrapperWithState<typename Integrand::template variant<V>, dim, test_space, trial_spaces...>;
  int variant_index() const { return integrand_.variant_index(); }
  template <int V>
  variant<V> as_variant() const
  {
    return variant<V>(integrand_.template as_variant<V>());
  }

  /// tangent fast path: argument i + 1 of the wrapper is argument i of the integrand; the shape correction only
  /// post-multiplies the linear maps (du/dx' = du/dX inv(J_p), flux -> flux inv(J_p)^T det J_p)
  template <int WRT>
  static constexpr bool has_flux_tangent = (WRT >= 1) && provides_flux_tangent<Integrand, WRT - 1>::value;
  /// (M -> M inv(J_p) and scale -> scale det J_p are congruences: a symmetric tangent stays symmetric)
  template <int WRT>
  static constexpr bool has_symmetric_flux_tangent = (WRT >= 1) && provides_symmetric_flux_tangent<Integrand, WRT - 1>::value;

  template <int WRT, typename PositionType, typename StateType, typename Source, typename Flux, typename Sink,
            typename ShapeValueType, typename... QFuncArgs>
  SERAC_HOST_DEVICE void flux_tangent(std::integral_constant<int, WRT>, double time, const PositionType& x,
                                      StateType& state, const ParentMap<dim>& map, Source& source, Flux& flux,
                                      Sink&& sink, const ShapeValueType& shape_val, const QFuncArgs&... qfunc_args) const
  {
    detail::ShapeCorrection<dim, ShapeValueType> shape_correction(Dimension<dim>{}, shape_val);
    ParentMap<dim> inner{dot(map.M, shape_correction.inv_J_), map.scale * shape_correction.detJ_};
    auto x_mod = serac::tuple{get<VALUE>(x) + get<VALUE>(shape_val),
                              get<DERIVATIVE>(x) + get<DERIVATIVE>(shape_val) * get<DERIVATIVE>(x)};
    integrand_.flux_tangent(std::integral_constant<int, WRT - 1>{}, time, x_mod, state, inner, source, flux, sink,
                            shape_correction.modify_trial_argument(trial_spaces{}, qfunc_args)...);
  }
};

// ------------------------------------------------------------------------------------------------
// boundary integrands (q-functions of AddBoundaryIntegral: `(t, tuple{x, dX/dxi}, tuple{value, d/dxi}...) -> source`)
// ------------------------------------------------------------------------------------------------
namespace detail {
/// ratio of the shape-displaced surface element to the reference one (shape_aware_functional.hpp:155-174)
template <typename position_type, typename shape_type>
SERAC_HOST_DEVICE auto compute_boundary_area_correction(const position_type& X, const shape_type& shape)
{
  auto x_prime = X + shape;
  auto n       = cross(get<DERIVATIVE>(x_prime));
  return norm(n) / norm(cross(get<DERIVATIVE>(X)));
}
}  // namespace detail

/// what ShapeAwareFunctional::AddBoundaryIntegral wraps every boundary integrand in
/// (shape_aware_functional.hpp:474-499,515-520): argument 0 is the shape displacement
template <typename Integrand>
struct ShapeAwareBoundaryIntegrandWrapper {
  SERAC_HOST_DEVICE ShapeAwareBoundaryIntegrandWrapper(Integrand integrand) : integrand_(integrand) {}
  Integrand integrand_;

  template <typename PositionType, typename ShapeValueType, typename... QFuncArgs>
  SERAC_HOST_DEVICE auto operator()(double time, PositionType x, ShapeValueType shape_val,
                                    QFuncArgs... qfunc_args) const
  {
    auto unmodified_qf_return = integrand_(time, x + shape_val, qfunc_args...);
    return unmodified_qf_return * detail::compute_boundary_area_correction(x, shape_val);
  }
};

/// the boundary q-function SolidMechanics::setTraction registers (solid_mechanics.hpp:1030-1036);
/// traction_function(x, n, t) -> traction vector
template <typename TractionType>
struct TractionIntegrand {
  SERAC_HOST_DEVICE TractionIntegrand(TractionType f) : traction_function(f) {}
  TractionType traction_function;

  template <typename X, typename Displacement, typename Acceleration, typename... Params>
  SERAC_HOST_DEVICE auto operator()(double t, X x, Displacement /*displacement*/, Acceleration /*acceleration*/,
                                    Params... params) const
  {
    auto n = cross(get<DERIVATIVE>(x));
    return -1.0 * traction_function(get<VALUE>(x), normalize(n), t, params...);
  }
};

/// the boundary q-function SolidMechanics::setPressure registers (solid_mechanics.hpp:1075-1101): follower load applied
/// on the current configuration when geometric nonlinearities are on; pressure_function(x, t) -> scalar
template <typename PressureType>
struct PressureIntegrand {
  SERAC_HOST_DEVICE PressureIntegrand(PressureType f, GeometricNonlinearities gn) : pressure_function(f), geom_nonlin(gn)
  {
  }
  PressureType            pressure_function;
  GeometricNonlinearities geom_nonlin;

  template <typename X, typename Displacement, typename Acceleration, typename... Params>
  SERAC_HOST_DEVICE auto operator()(double t, X Xref, Displacement displacement, Acceleration /*acceleration*/,
                                    Params... params) const
  {
    auto x = Xref + 0.0 * displacement;
    if (geom_nonlin == GeometricNonlinearities::On) {
      x = x + displacement;
    }
    auto n = cross(get<DERIVATIVE>(x));
    return pressure_function(get<VALUE>(Xref), t, params...) * (n / norm(cross(get<DERIVATIVE>(Xref))));
  }
};

/// load functions for the named integrands: constant traction vector / constant pressure
template <int dim>
struct UniformTraction {
  tensor<double, dim> traction;
  template <typename X, typename N>
  SERAC_HOST_DEVICE tensor<double, dim> operator()(const X&, const N&, double) const
  {
    return traction;
  }
};
struct UniformPressure {
  double pressure;
  template <typename X>
  SERAC_HOST_DEVICE double operator()(const X&, double) const
  {
    return pressure;
  }
};

}  // namespace serac
