// TEST INFRASTRUCTURE ONLY (oracle/): instantiations of the REFERENCE's MakeDomainIntegral for the q-functions the
// physics modules register.  Compiled once per (REF_DIM, REF_P).  The two functor shells below restate, with the
// reference's own tensor/tuple/dual types, what lives inside class templates that need MPI to instantiate:
//   SolidMechanics::MaterialStressFunctor       src/serac/physics/solid_mechanics.hpp:808-853
//   ShapeAwareFunctional::ShapeAwareIntegrandWrapper[WithState]  shape_aware_functional.hpp:364-438
// (the shape pull-backs themselves -- detail::ShapeCorrection and apply_shape_aware_qf_helper* -- are the reference's).
#include "ref_driver.hpp"

#ifndef REF_DIM
#error "compile with -DREF_DIM=2|3 -DREF_P=1|2|3"
#endif

namespace {
using namespace serac;
constexpr int dim = REF_DIM;
constexpr int p   = REF_P;

template <typename Material>
struct MaterialStress {
  Material material_;
  bool     geom_nonlin_;
  template <typename X, typename State, typename Displacement, typename Acceleration>
  auto operator()(double, X, State& state, Displacement displacement, Acceleration acceleration) const
  {
    auto du_dX   = get<1>(displacement);
    auto d2u_dt2 = get<0>(acceleration);
    auto stress  = material_(state, du_dX);
    auto dx_dX   = 0.0 * du_dX + Identity<dim>();
    if (geom_nonlin_) dx_dX += du_dX;
    auto flux = dot(stress, transpose(inv(dx_dX))) * det(dx_dX);
    return serac::tuple{material_.density * d2u_dt2, flux};
  }
};

template <typename Material>
struct ThermalIntegrand {
  Material material_;
  template <typename X, typename T, typename dT_dt>
  auto operator()(double, X x, T temperature, dT_dt dtemp_dt) const
  {
    auto [u, du_dX]                 = temperature;
    auto du_dt                      = get<0>(dtemp_dt);
    auto [heat_capacity, heat_flux] = material_(x, u, du_dX);
    return serac::tuple{heat_capacity * du_dt, -1.0 * heat_flux};
  }
};

template <typename Integrand, typename test_space, typename... trial_spaces>
struct ShapeAwareWithState {
  Integrand integrand_;
  template <typename Pos, typename State, typename Shape, typename... Args>
  auto operator()(double time, Pos x, State& state, Shape shape_val, Args... args) const
  {
    auto                    qfunc_tuple = make_tuple(args...);
    auto                    spaces      = make_tuple(trial_spaces{}...);
    serac::detail::ShapeCorrection shape_correction(Dimension<dim>{}, shape_val);
    auto unmodified = serac::detail::apply_shape_aware_qf_helper_with_state(
        integrand_, time, x, state, shape_val, spaces, qfunc_tuple, shape_correction,
        std::make_integer_sequence<int, sizeof...(args)>{});
    return shape_correction.modify_shape_aware_qf_return(test_space{}, unmodified);
  }
};

template <typename Integrand, typename test_space, typename... trial_spaces>
struct ShapeAware {
  Integrand integrand_;
  template <typename Pos, typename Shape, typename... Args>
  auto operator()(double time, Pos x, Shape shape_val, Args... args) const
  {
    auto                    qfunc_tuple = make_tuple(args...);
    auto                    spaces      = make_tuple(trial_spaces{}...);
    serac::detail::ShapeCorrection shape_correction(Dimension<dim>{}, shape_val);
    auto unmodified = serac::detail::apply_shape_aware_qf_helper(integrand_, time, x, shape_val, spaces, qfunc_tuple,
                                                          shape_correction,
                                                          std::make_integer_sequence<int, sizeof...(args)>{});
    return shape_correction.modify_shape_aware_qf_return(test_space{}, unmodified);
  }
};

constexpr mfem::Geometry::Type tp_geom = (dim == 3) ? mfem::Geometry::CUBE : mfem::Geometry::SQUARE;

template <typename State>
std::shared_ptr<QuadratureData<State>> make_qdata(mfem::Mesh& mesh, ref::IntegralHolder& h, State init)
{
  // StateManager::newQuadratureDataBuffer (state_manager.hpp:108-132): elements per geometry x qpts per element
  typename QuadratureData<State>::geom_array_t elems{}, qpts{};
  for (int e = 0; e < mesh.GetNE(); e++) elems[size_t(mesh.GetElementGeometry(e))]++;
  for (int g = 0; g < mfem::Geometry::NUM_GEOMETRIES; g++)
    qpts[size_t(g)] = uint32_t(std::max(0, num_quadrature_points(mfem::Geometry::Type(g), p + 1)));
  auto qd        = std::make_shared<QuadratureData<State>>(elems, qpts, init);
  h.qdata        = qd;
  h.state_bytes  = sizeof(State);
  // single-geometry meshes: the buffer of the geometry the mesh is made of (hex / quad / tet / tri)
  const auto g   = (mesh.GetNE() > 0) ? mesh.GetElementGeometry(0) : tp_geom;
  h.qdata_count  = size_t(elems[size_t(g)]) * size_t(qpts[size_t(g)]);
  h.qdata_raw    = (h.qdata_count > 0) ? static_cast<void*>(&(*qd)[g](0, 0)) : nullptr;
  return qd;
}

using solid_test  = H1<p, dim>;
using solid_sig   = solid_test(H1<1, dim>, H1<p, dim>, H1<p, dim>);
using thermal_sig = H1<p>(H1<1, dim>, H1<p>, H1<p>);

template <typename Material>
ref::IntegralHolder make_solid_stateless(mfem::Mesh& mesh, Material m, bool gn)
{
  ref::IntegralHolder h;
  using QF = ShapeAwareWithState<MaterialStress<Material>, solid_test, H1<p, dim>, H1<p, dim>>;
  QF     qf{MaterialStress<Material>{m, gn}};
  Domain d     = ref::domain_for(mesh);
  h.integral   = std::make_unique<Integral>(MakeDomainIntegral<solid_sig, p + 1, dim>(d, qf, EmptyQData, {0, 1, 2}));
  h.num_trials = 3;
  return h;
}

#define STR2(x) #x
#define STR(x) STR2(x)
#define TAG STR(REF_DIM) "d_p" STR(REF_P)

ref::Registrar r1("solid_neohookean_" TAG, [](mfem::Mesh& mesh, const double* q, int n) {
  SLIC_ERROR_IF(n != 4, "params: density K G geom_nonlin");
  return make_solid_stateless(mesh, solid_mechanics::NeoHookean{q[0], q[1], q[2]}, q[3] != 0.0);
});
ref::Registrar r2("solid_linear_isotropic_" TAG, [](mfem::Mesh& mesh, const double* q, int n) {
  SLIC_ERROR_IF(n != 4, "params: density K G geom_nonlin");
  return make_solid_stateless(mesh, solid_mechanics::LinearIsotropic{q[0], q[1], q[2]}, q[3] != 0.0);
});

#if REF_DIM == 3
ref::Registrar r3("solid_j2smallstrain_linear_" TAG, [](mfem::Mesh& mesh, const double* q, int n) {
  SLIC_ERROR_IF(n != 7, "params: E nu sigma_y Hi Hk density geom_nonlin");
  using Material = solid_mechanics::J2SmallStrain<solid_mechanics::LinearHardening>;
  Material m{q[0], q[1], solid_mechanics::LinearHardening{q[2], q[3]}, q[4], q[5]};
  ref::IntegralHolder h;
  using QF = ShapeAwareWithState<MaterialStress<Material>, solid_test, H1<p, dim>, H1<p, dim>>;
  QF     qf{MaterialStress<Material>{m, q[6] != 0.0}};
  Domain d  = ref::domain_for(mesh);
  auto   qd = make_qdata<Material::State>(mesh, h, Material::State{});
  h.integral   = std::make_unique<Integral>(MakeDomainIntegral<solid_sig, p + 1, dim>(d, qf, qd, {0, 1, 2}));
  h.num_trials = 3;
  return h;
});
#endif

#if REF_DIM == 3
// the other hardening laws of solid_material.hpp (:154-198) under the same small-strain J2 model
template <typename Hardening>
ref::IntegralHolder make_j2_small(mfem::Mesh& mesh, Hardening hardening, const double* q)
{
  using Material = solid_mechanics::J2SmallStrain<Hardening>;
  Material m{q[0], q[1], hardening, q[5], q[6]};
  ref::IntegralHolder h;
  using QF = ShapeAwareWithState<MaterialStress<Material>, solid_test, H1<p, dim>, H1<p, dim>>;
  QF     qf{MaterialStress<Material>{m, q[7] != 0.0}};
  Domain d  = ref::domain_for(mesh);
  auto   qd = make_qdata<typename Material::State>(mesh, h, typename Material::State{});
  h.integral   = std::make_unique<Integral>(MakeDomainIntegral<solid_sig, p + 1, dim>(d, qf, qd, {0, 1, 2}));
  h.num_trials = 3;
  return h;
}
ref::Registrar r3p("solid_j2smallstrain_powerlaw_" TAG, [](mfem::Mesh& mesh, const double* q, int n) {
  SLIC_ERROR_IF(n != 8, "params: E nu sigma_y n eps0 Hk density geom_nonlin");
  return make_j2_small(mesh, solid_mechanics::PowerLawHardening{q[2], q[3], q[4]}, q);
});
ref::Registrar r3v("solid_j2smallstrain_voce_" TAG, [](mfem::Mesh& mesh, const double* q, int n) {
  SLIC_ERROR_IF(n != 8, "params: E nu sigma_y sigma_sat strain_constant Hk density geom_nonlin");
  return make_j2_small(mesh, solid_mechanics::VoceHardening{q[2], q[3], q[4]}, q);
});
ref::Registrar r3s("solid_stvk_" TAG, [](mfem::Mesh& mesh, const double* q, int n) {
  SLIC_ERROR_IF(n != 4, "params: density K G geom_nonlin");
  return make_solid_stateless(mesh, solid_mechanics::StVenantKirchhoff{q[0], q[1], q[2]}, q[3] != 0.0);
});
#endif

#if REF_DIM == 3
ref::Registrar r3f("solid_j2_linear_" TAG, [](mfem::Mesh& mesh, const double* q, int n) {
  SLIC_ERROR_IF(n != 6, "params: E nu sigma_y Hi density geom_nonlin");
  using Material = solid_mechanics::J2<solid_mechanics::LinearHardening>;
  Material m{q[0], q[1], solid_mechanics::LinearHardening{q[2], q[3]}, q[4]};
  ref::IntegralHolder h;
  using QF = ShapeAwareWithState<MaterialStress<Material>, solid_test, H1<p, dim>, H1<p, dim>>;
  QF     qf{MaterialStress<Material>{m, q[5] != 0.0}};
  Domain d  = ref::domain_for(mesh);
  auto   qd = make_qdata<Material::State>(mesh, h, Material::State{});  // Fpinv = I, eqps = 0 (solid_material.hpp:272-275)
  h.integral   = std::make_unique<Integral>(MakeDomainIntegral<solid_sig, p + 1, dim>(d, qf, qd, {0, 1, 2}));
  h.num_trials = 3;
  return h;
});
#endif

ref::Registrar r4("thermal_linear_conductor_" TAG, [](mfem::Mesh& mesh, const double* q, int n) {
  SLIC_ERROR_IF(n != 3, "params: density specific_heat conductivity");
  using Material = heat_transfer::LinearIsotropicConductor;
  ref::IntegralHolder h;
  using QF = ShapeAware<ThermalIntegrand<Material>, H1<p>, H1<p>, H1<p>>;
  QF     qf{ThermalIntegrand<Material>{Material(q[0], q[1], q[2])}};
  Domain d     = ref::domain_for(mesh);
  h.integral   = std::make_unique<Integral>(MakeDomainIntegral<thermal_sig, p + 1, dim>(d, qf, NoQData, {0, 1, 2}));
  h.num_trials = 3;
  return h;
});

// ---- a discontinuous (L2) parameter field: H1<p>(H1<p>, L2<0>), reaction-diffusion with a piecewise-constant coefficient
// (how the physics modules take material parameters, e.g. Parameters<L2<0>> in solid_mechanics.hpp)
struct ParameterizedDiffusion {
  double c0;
  template <typename X, typename U, typename K>
  auto operator()(double, X, U u, K k) const
  {
    auto [uv, du] = u;
    auto kv       = get<0>(k);
    return serac::tuple{(c0 + kv) * uv, (kv * kv) * du};
  }
};
ref::Registrar rp("param_diffusion_" TAG, [](mfem::Mesh& mesh, const double* q, int n) {
  SLIC_ERROR_IF(n != 1, "params: c0");
  ref::IntegralHolder h;
  using sig    = H1<p>(H1<p>, L2<0>);
  Domain d     = ref::domain_for(mesh);
  h.integral   = std::make_unique<Integral>(MakeDomainIntegral<sig, p + 1, dim>(d, ParameterizedDiffusion{q[0]}, NoQData, {0, 1}));
  h.num_trials = 2;
  return h;
});

// ---- quantity of interest: Functional<double(H1<p,dim>)> (functional_qoi.inl:160-190 builds signature QOI(trials...)) ----
struct QuadraticQoi {
  double a, b;
  template <typename X, typename U>
  auto operator()(double, X x, U u) const
  {
    using std::sin;
    auto [uv, du] = u;
    auto X0       = get<0>(x);
    return a * (1.0 + X0[0]) * dot(uv, uv) + (0.5 * b) * inner(du, du) + sin(X0[1]);
  }
};

ref::Registrar rq("qoi_quadratic_" TAG, [](mfem::Mesh& mesh, const double* q, int n) {
  SLIC_ERROR_IF(n != 2, "params: a b");
  ref::IntegralHolder h;
  using sig    = QOI(H1<p, dim>);
  Domain d     = ref::domain_for(mesh);
  h.integral   = std::make_unique<Integral>(MakeDomainIntegral<sig, p + 1, dim>(d, QuadraticQoi{q[0], q[1]}, NoQData, {0}));
  h.num_trials = 1;
  return h;
});

#if REF_DIM == 3
// ---- thermomechanics (BASELINE.json configs[3]) ---------------------------------------------------------------------
// functor shells restating Thermomechanics::ThermalMaterialInterface / MechanicalMaterialInterface
// (src/serac/physics/thermomechanics.hpp:272-352), which live inside a class template that needs MPI to instantiate;
// the material itself (materials/green_saint_venant_thermoelastic.hpp) and the integrand functors above are the
// reference's, with the parameter pack forwarded as SolidMechanics / HeatTransfer do (solid_mechanics.hpp:834-852,
// heat_transfer.hpp:392-402).
template <typename Material>
struct MechanicalInterface {
  using State = typename Material::State;
  Material mat;
  double   density;
  template <typename T1, typename T2>
  auto operator()(State& state, const T1& displacement_gradient, const T2& temperature) const
  {
    auto [theta, dtheta_dX]         = temperature;
    auto [T, heat_capacity, s0, q0] = mat(state, displacement_gradient, theta, dtheta_dX);
    return T;
  }
};

template <typename Material>
struct ThermalInterface {
  using State = typename Material::State;
  Material mat;
  template <typename T1, typename T2, typename T3, typename T4>
  auto operator()(const T1&, const T2& temperature, const T3& temperature_gradient, const T4& displacement) const
  {
    State state{};
    auto [u, du_dX]                 = displacement;
    auto [T, heat_capacity, s0, q0] = mat(state, du_dX, temperature, temperature_gradient);
    return serac::tuple{heat_capacity, q0};
  }
};

template <typename Material>
struct MaterialStressWithParam {
  Material material_;
  bool     geom_nonlin_;
  template <typename X, typename State, typename Displacement, typename Acceleration, typename Param>
  auto operator()(double, X, State& state, Displacement displacement, Acceleration acceleration, Param param) const
  {
    auto du_dX   = get<1>(displacement);
    auto d2u_dt2 = get<0>(acceleration);
    auto stress  = material_(state, du_dX, param);
    auto dx_dX   = 0.0 * du_dX + Identity<dim>();
    if (geom_nonlin_) dx_dX += du_dX;
    auto flux = dot(stress, transpose(inv(dx_dX))) * det(dx_dX);
    return serac::tuple{material_.density * d2u_dt2, flux};
  }
};

template <typename Material>
struct ThermalIntegrandWithParam {
  Material material_;
  template <typename X, typename T, typename dT_dt, typename Param>
  auto operator()(double, X x, T temperature, dT_dt dtemp_dt, Param param) const
  {
    auto [u, du_dX]                 = temperature;
    auto du_dt                      = get<0>(dtemp_dt);
    auto [heat_capacity, heat_flux] = material_(x, u, du_dX, param);
    return serac::tuple{heat_capacity * du_dt, -1.0 * heat_flux};
  }
};

using GSV = GreenSaintVenantThermoelasticMaterial;

ref::Registrar r7("thermoelastic_solid_" TAG, [](mfem::Mesh& mesh, const double* q, int n) {
  SLIC_ERROR_IF(n != 8, "params: density E nu C_v alpha theta_ref k geom_nonlin");
  GSV mat{q[0], q[1], q[2], q[3], q[4], q[5], q[6]};
  using Mech = MechanicalInterface<GSV>;
  using sig  = solid_test(H1<1, dim>, H1<p, dim>, H1<p, dim>, H1<p>);
  using QF   = ShapeAwareWithState<MaterialStressWithParam<Mech>, solid_test, H1<p, dim>, H1<p, dim>, H1<p>>;
  QF     qf{MaterialStressWithParam<Mech>{Mech{mat, mat.density}, q[7] != 0.0}};
  ref::IntegralHolder h;
  Domain d  = ref::domain_for(mesh);
  auto   qd = make_qdata<GSV::State>(mesh, h, GSV::State{});
  h.integral   = std::make_unique<Integral>(MakeDomainIntegral<sig, p + 1, dim>(d, qf, qd, {0, 1, 2, 3}));
  h.num_trials = 4;
  return h;
});

ref::Registrar r8("thermoelastic_thermal_" TAG, [](mfem::Mesh& mesh, const double* q, int n) {
  SLIC_ERROR_IF(n != 7, "params: density E nu C_v alpha theta_ref k");
  GSV mat{q[0], q[1], q[2], q[3], q[4], q[5], q[6]};
  using Therm = ThermalInterface<GSV>;
  using sig   = H1<p>(H1<1, dim>, H1<p>, H1<p>, H1<p, dim>);
  using QF    = ShapeAware<ThermalIntegrandWithParam<Therm>, H1<p>, H1<p>, H1<p>, H1<p, dim>>;
  QF     qf{ThermalIntegrandWithParam<Therm>{Therm{mat}}};
  ref::IntegralHolder h;
  Domain d     = ref::domain_for(mesh);
  h.integral   = std::make_unique<Integral>(MakeDomainIntegral<sig, p + 1, dim>(d, qf, NoQData, {0, 1, 2, 3}));
  h.num_trials = 4;
  return h;
});

#endif  // REF_DIM == 3

// ---- boundary integrands (quadrilateral / triangular faces of 3D meshes, segment faces of 2D meshes) -------------------------------------------------------------------
// functor shells restating ShapeAwareFunctional::ShapeAwareBoundaryIntegrandWrapper (shape_aware_functional.hpp:474-499)
// and the lambdas of SolidMechanics::setTraction / setPressure (solid_mechanics.hpp:1030-1036, 1075-1101) with a
// constant traction vector / pressure; compute_boundary_area_correction is the reference's.
template <typename Integrand>
struct ShapeAwareBoundary {
  Integrand integrand_;
  template <typename Pos, typename Shape, typename... Args>
  auto operator()(double time, Pos x, Shape shape_val, Args... args) const
  {
    auto unmodified_qf_return = integrand_(time, x + shape_val, args...);
    return unmodified_qf_return * serac::detail::compute_boundary_area_correction(x, shape_val);
  }
};

struct UniformTractionIntegrand {
  tensor<double, dim> traction_;
  template <typename X, typename U, typename A>
  auto operator()(double, X x, U, A) const
  {
    auto n = cross(get<DERIVATIVE>(x));
    (void)normalize(n);  // the reference hands normalize(n) to the traction function, which ignores it here
    return -1.0 * traction_;
  }
};

struct UniformPressureIntegrand {
  double pressure_;
  bool   geom_nonlin_;
  template <typename X, typename U, typename A>
  auto operator()(double, X Xref, U displacement, A) const
  {
    auto x = Xref + 0.0 * displacement;
    if (geom_nonlin_) {
      x = x + displacement;
    }
    auto n = cross(get<DERIVATIVE>(x));
    return pressure_ * (n / norm(cross(get<DERIVATIVE>(Xref))));
  }
};

template <typename QF>
ref::IntegralHolder make_boundary(mfem::Mesh& mesh, QF qf)
{
  ref::IntegralHolder h;
  Domain d     = EntireBoundary(mesh);
  h.integral   = std::make_unique<Integral>(MakeBoundaryIntegral<solid_sig, p + 1, dim - 1>(d, qf, {0, 1, 2}));
  h.num_trials = 3;
  return h;
}

// boundary quantity of interest (functional_qoi.inl:196-215: MakeBoundaryIntegral<QOI(trials...)>)
struct QuadraticBoundaryQoi {
  double a, b;
  template <typename X, typename U>
  auto operator()(double, X x, U u) const
  {
    auto [uv, du] = u;
    auto X0       = get<0>(x);
    return a * (1.0 + X0[0]) * dot(uv, uv) + (0.5 * b) * inner(du, du);
  }
};
ref::Registrar rqb("qoi_boundary_quadratic_" TAG, [](mfem::Mesh& mesh, const double* q, int n) {
  SLIC_ERROR_IF(n != 2, "params: a b");
  ref::IntegralHolder h;
  using sig    = QOI(H1<p, dim>);
  Domain d     = EntireBoundary(mesh);
  h.integral   = std::make_unique<Integral>(MakeBoundaryIntegral<sig, p + 1, dim - 1>(d, QuadraticBoundaryQoi{q[0], q[1]}, {0}));
  h.num_trials = 1;
  return h;
});

ref::Registrar r5("solid_traction_" TAG, [](mfem::Mesh& mesh, const double* q, int n) {
  SLIC_ERROR_IF(n != dim, "params: tx ty [tz]");
  tensor<double, dim> traction{};
  for (int i = 0; i < dim; i++) traction[i] = q[i];
  return make_boundary(mesh, ShapeAwareBoundary<UniformTractionIntegrand>{UniformTractionIntegrand{traction}});
});
ref::Registrar r6("solid_pressure_" TAG, [](mfem::Mesh& mesh, const double* q, int n) {
  SLIC_ERROR_IF(n != 2, "params: pressure geom_nonlin");
  return make_boundary(mesh, ShapeAwareBoundary<UniformPressureIntegrand>{UniformPressureIntegrand{q[0], q[1] != 0.0}});
});

}  // namespace
