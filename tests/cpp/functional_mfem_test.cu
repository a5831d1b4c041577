// Test driver for the mfem-typed boundary (serac_b200/functional_mfem.hpp): the reference's own unit tests of Functional,
// ported with minimal edits, running through `serac::Functional` on mfem::FiniteElementSpace / mfem::Vector /
// mfem::Operator, linked against the reference's serial mfem (oracle/_ref/mfem-install, tests only).
//
//   comparisons   src/serac/numerics/functional/tests/functional_comparisons.cpp:78-330,431-445: residual, Jacobian action
//                 and assembled matrix of a toy thermal problem (H1<p>) and a toy elasticity problem (H1<p,dim>) against
//                 MFEM's own integrators (MassIntegrator + DiffusionIntegrator, VectorMassIntegrator +
//                 ElasticityIntegrator), tolerance 1e-14 relative as in the reference (:283, :314)
//   gradients     functional_basic_h1_vector.cpp:60-165 + check_gradient.hpp:16-90: domain + boundary integrals with
//                 source / flux terms of every shape; JVP against the assembled matrix (5e-6) and finite-difference
//                 convergence ratios
//   scalar        functional_basic_h1_scalar.cpp:32-118: H1<ptest>(H1<ptrial>) with a domain and a boundary integral on
//                 hexahedra, tetrahedra, quadrilaterals and triangles (boundary faces of the 2D meshes: segments),
//                 equal and mixed orders, through check_gradient
//   domains       functional_with_domain.cpp:89-316: Domain::ofElements / ofBoundaryElements with the reference's predicates;
//                 integrals on complementary parts add up to the whole-mesh Functional (every geometry incl. mixed
//                 meshes), a domain equals a masked integrand, quantities of interest on part of the mesh / boundary
//   boundary      functional_boundary_test.cpp:59-233: boundary integrals of H1 and of discontinuous L2 spaces (face dofs of a
//                 DG space = those of the adjacent element, element_restriction.cpp:304-355) against MFEM's
//                 BoundaryLFIntegrator / BoundaryNormalLFIntegrator / BoundaryMassIntegrator on the mixed meshes;
//                 bug_boundary_qoi.cpp: ShapeAwareFunctional<shape, double()> on the faces of one boundary attribute
//   nonlinear     functional_nonlinear.cpp:45-153 (thermal and elasticity, p = 1, 2, 3, on the mixed 2D and 3D meshes) and
//                 functional_multiphysics.cpp:25-83 (two arguments, p = 3, AddVolumeIntegral + AddSurfaceIntegral) through
//                 check_gradient
//   qoi           functional_qoi.cpp:196-452: Functional<double(...)> -- measures, moments, volume + surface against closed
//                 forms and MFEM's linear forms, nonlinear and two-argument QOIs through check_gradient's QOI overloads
//                 (central difference vs assembled gradient vector vs directional derivative), vector-valued and L2
//                 arguments, AddAreaIntegral / AddVolumeIntegral
//   shape         functional_shape_derivatives.cpp: the divergence theorem on the shape-displaced mixed meshes (the residual
//                 and its derivative w.r.t. the shape field sum to zero);
//                 ShapeAwareFunctional (shape_aware_functional.hpp:292-575): with a shape displacement p the residual
//                 equals that of the plain Functional on the mesh whose nodes were moved by p
// Meshes are generated (MakeCartesian3D + a smooth perturbation): the reference's mesh files are not on the GPU box.
// Exit code 0 = all checks passed; every check prints one line.
#include <cmath>
#include <cstdio>
#include <string>

#include "serac_b200/functional_mfem.hpp"

using namespace serac;

static int g_failures = 0;
#define EXPECT(cond, ...)                                  \
  do {                                                     \
    const bool _ok = (cond);                               \
    std::printf("%s  ", _ok ? "ok  " : "FAIL");            \
    std::printf(__VA_ARGS__);                              \
    std::printf("\n");                                     \
    if (!_ok) g_failures++;                                \
  } while (0)

// (literals in device code: a namespace-scope constexpr double that is passed by reference is not usable on the device)
#define COEF_A 1.7
#define COEF_B 2.1
constexpr double a = COEF_A;
constexpr double b = COEF_B;

// ---- q-functions of the reference tests (functor structs, as in the reference's CUDA-enabled tests) ---------------------
struct SourceFluxFunctor {  // functional_comparisons.cpp:67-76
  template <typename Position, typename Temperature>
  SERAC_HOST_DEVICE auto operator()(double /*t*/, Position position, Temperature temperature) const
  {
    auto [X, dX_dxi] = position;
    auto [u, du_dX]  = temperature;
    auto source      = COEF_A * u - (100 * X[0] * X[1]);
    auto flux        = COEF_B * du_dX;
    return serac::tuple{source, flux};
  }
};

template <int dim>
struct StressFunctor {  // functional_comparisons.cpp:53-65
  template <typename Position, typename Displacement>
  SERAC_HOST_DEVICE auto operator()(double /*t*/, Position, Displacement displacement) const
  {
    const auto I    = DenseIdentity<dim>();
    auto [u, du_dx] = displacement;
    auto body_force = COEF_A * u + I[0];
    auto strain     = 0.5 * (du_dx + transpose(du_dx));
    auto stress     = COEF_B * tr(strain) * I + 2.0 * COEF_B * strain;
    return serac::tuple{body_force, stress};
  }
};

template <int dim>
struct ElasticityTestModelOne {  // functional_basic_h1_vector.cpp:60-75
  template <typename position_type, typename displacement_type>
  SERAC_HOST_DEVICE auto operator()(double, position_type, displacement_type displacement) const
  {
    const auto d00 = make_tensor<dim, dim>([](int i, int j) { return i + 2 * j + 1; });
    const auto d01 = make_tensor<dim, dim, dim>([](int i, int j, int k) { return i + 2 * j - k + 1; });
    const auto d10 = make_tensor<dim, dim, dim>([](int i, int j, int k) { return i + 3 * j - 2 * k; });
    const auto d11 = make_tensor<dim, dim, dim, dim>([](int i, int j, int k, int l) { return i - j + 2 * k - 3 * l + 1; });
    auto [u, du_dx] = displacement;
    auto source     = dot(d00, u) + double_dot(d01, du_dx);
    auto flux       = dot(d10, u) + double_dot(d11, du_dx);
    return serac::tuple{source, flux};
  }
};

template <int dim>
struct ElasticityTestModelTwo {  // functional_basic_h1_vector.cpp:77-87 (boundary integrand)
  template <typename position_type, typename displacement_type>
  SERAC_HOST_DEVICE auto operator()(double, position_type position, displacement_type displacement) const
  {
    auto [X, dX_dxi] = position;
    auto [u, du_dxi] = displacement;
    return u * X[0];
  }
};

template <int dim>
struct TestThermalModelOne {  // functional_basic_h1_scalar.cpp:32-49
  template <typename P, typename Temp>
  SERAC_HOST_DEVICE auto operator()(double, P position, Temp temperature) const
  {
    double     d00 = 1.0;
    const auto d01 = 1.0 * make_tensor<dim>([](int i) { return i; });
    const auto d10 = 1.0 * make_tensor<dim>([](int i) { return 2 * i * i; });
    const auto d11 = 1.0 * make_tensor<dim, dim>([](int i, int j) { return i + j * (j + 1) + 1; });
    auto [X, dX_dxi] = position;
    auto [u, du_dx]  = temperature;
    auto source      = d00 * u + dot(d01, du_dx) - 0.0 * (100 * X[0] * X[1]);
    auto flux        = d10 * u + dot(d11, du_dx);
    return serac::tuple{source, flux};
  }
};

struct TestThermalModelTwo {  // functional_basic_h1_scalar.cpp:51-60 (boundary integrand)
  template <typename PositionType, typename TempType>
  SERAC_HOST_DEVICE auto operator()(double, PositionType position, TempType temperature) const
  {
    auto [X, dX_dxi] = position;
    auto [u, du_dxi] = temperature;
    return X[0] + X[1] - cos(u);
  }
};

template <int dim, bool apply_mask = false>
struct TestThermalIntegratorOne {  // functional_with_domain.cpp:26-50
  template <typename Position, typename Temperature>
  SERAC_HOST_DEVICE auto operator()(double, Position position, Temperature temperature) const
  {
    const double d00 = 1.0;
    const auto   d01 = 1.0 * make_tensor<dim>([](int i) { return i; });
    const auto   d10 = 1.0 * make_tensor<dim>([](int i) { return 2 * i * i; });
    const auto   d11 = 1.0 * make_tensor<dim, dim>([](int i, int j) { return i + j * (j + 1) + 1; });
    auto [X, dX_dxi] = position;
    auto [u, du_dx]  = temperature;
    double mask = 1.0;
    if constexpr (apply_mask) {
      mask = (X[0] < 4.0);
    }
    auto source = mask * (d00 * u + dot(d01, du_dx) - 0.0 * (100 * X[0] * X[1]));
    auto flux   = mask * (d10 * u + dot(d11, du_dx));
    return serac::tuple{source, flux};
  }
};

template <bool apply_mask = false>
struct TestThermalIntegratorTwo {  // functional_with_domain.cpp:52-68
  template <typename Position, typename Temperature>
  SERAC_HOST_DEVICE auto operator()(double, Position position, Temperature temperature) const
  {
    auto [X, dX_dxi] = position;
    auto [u, du_dxi] = temperature;
    double mask = 1.0;
    if constexpr (apply_mask) {
      mask = (X[1] >= 0.99);
    }
    return mask * (X[0] + X[1] - cos(u));
  }
};

struct TrivialIntegrator {  // functional_with_domain.cpp:70-77
  template <typename UnusedType>
  SERAC_HOST_DEVICE auto operator()(double /*t*/, UnusedType /*position*/) const
  {
    return 1.0;
  }
};

template <int dim>
tensor<double, dim> average(std::vector<tensor<double, dim>>& positions)  // functional_with_domain.cpp:79-87
{
  tensor<double, dim> total{};
  for (auto x : positions) total += x;
  return total / double(positions.size());
}

// ---- functional_qoi.cpp:35-128 ------------------------------------------------------------------------------------------
struct TrivialVariadicIntegrator {
  template <typename... Args>
  SERAC_HOST_DEVICE auto operator()(Args...) const
  {
    return 1.0;
  }
};
struct ZeroIndexIntegrator {
  template <typename P>
  SERAC_HOST_DEVICE auto operator()(double /*t*/, P position) const
  {
    return get<VALUE>(position)[0];
  }
};
struct GetZeroIntegrator {
  template <typename Unused, typename X>
  SERAC_HOST_DEVICE auto operator()(double /*t*/, Unused, X x) const
  {
    return serac::get<0>(x);
  }
};
struct SineIntegrator {
  template <typename Position, typename Temperature>
  SERAC_HOST_DEVICE auto operator()(double /*t*/, Position position, Temperature temperature) const
  {
    auto X           = get<VALUE>(position);
    auto [u, grad_u] = temperature;
    return X[0] * X[0] + sin(X[1]) + X[0] * u * u * u;
  }
};
struct CosineIntegrator {
  template <typename Position, typename Temperature>
  SERAC_HOST_DEVICE auto operator()(double /*t*/, Position position, Temperature temperature) const
  {
    auto [X, dX_dxi] = position;
    auto [u, unused] = temperature;
    return X[0] - X[1] + cos(u * X[1]);
  }
};
struct FourArgSineIntegrator {
  template <typename Position, typename Temperature, typename TimeDerivativeTemp>
  SERAC_HOST_DEVICE auto operator()(double /*t*/, Position position, Temperature temperature,
                                    TimeDerivativeTemp dtemperature_dt) const
  {
    auto [X, dX_dxi]     = position;
    auto [u, grad_u]     = temperature;
    auto [du_dt, unused] = dtemperature_dt;
    return X[0] * X[0] + sin(du_dt) + X[0] * u * u * u;
  }
};
struct FourArgCosineIntegrator {
  template <typename Position, typename Temperature, typename TimeDerivativeTemp>
  SERAC_HOST_DEVICE auto operator()(double /*t*/, Position position, Temperature temperature,
                                    TimeDerivativeTemp dtemperature_dt) const
  {
    auto [X, dX_dxi]     = position;
    auto [u, grad_u]     = temperature;
    auto [du_dt, unused] = dtemperature_dt;
    return X[0] - X[1] + cos(u * du_dt);
  }
};
struct GetNormZeroIntegrator {
  template <typename Unused, typename X>
  SERAC_HOST_DEVICE auto operator()(double /*t*/, Unused, X x) const
  {
    return norm(serac::get<0>(x));
  }
};

// ---- functional_nonlinear.cpp:57-66,74-79,106-113,121-127 (the reference's lambdas as functors) ---------------------------
struct NonlinearThermalDomain {
  template <typename Position, typename Temperature>
  SERAC_HOST_DEVICE auto operator()(double /*t*/, Position position, Temperature temperature) const
  {
    auto [X, dX_dxi] = position;
    auto [u, du_dX]  = temperature;
    auto source      = COEF_A * u * u - (100 * X[0] * X[1]);
    auto flux        = COEF_B * du_dX;
    return serac::tuple{source, flux};
  }
};
struct NonlinearThermalBoundary {
  template <typename Position, typename Temperature>
  SERAC_HOST_DEVICE auto operator()(double /*t*/, Position position, Temperature temperature) const
  {
    auto [X, dX_dxi] = position;
    auto u           = get<0>(temperature);
    return X[0] + X[1] - cos(u);
  }
};
struct NonlinearElasticDomain {
  template <typename Position, typename Displacement>
  SERAC_HOST_DEVICE auto operator()(double /*t*/, Position /*x*/, Displacement displacement) const
  {
    auto [u, du_dx] = displacement;
    auto source     = COEF_A * u * u[0];
    auto flux       = COEF_B * du_dx;
    return serac::tuple{source, flux};
  }
};
struct NonlinearElasticBoundary {
  template <typename Position, typename Displacement>
  SERAC_HOST_DEVICE auto operator()(double /*t*/, Position position, Displacement displacement) const
  {
    auto [X, dX_dxi] = position;
    auto u           = get<0>(displacement);
    auto n           = normalize(cross(dX_dxi));
    return (X[0] + X[1] - cos(u[0])) * n;
  }
};
// ---- functional_multiphysics.cpp:55-76 -------------------------------------------------------------------------------------
struct MultiphysicsVolume {
  template <typename Position, typename Temperature, typename Rate>
  SERAC_HOST_DEVICE auto operator()(double /*t*/, Position position, Temperature temperature, Rate dtemperature_dt) const
  {
    auto [X, dX_dxi]     = position;
    auto [u, du_dX]      = temperature;
    auto [du_dt, unused] = dtemperature_dt;
    auto source          = 1.0 * 1.0 * du_dt * du_dt - (100 * X[0] * X[1]);
    auto flux            = 1.0 * du_dX;
    return serac::tuple{source, flux};
  }
};
struct MultiphysicsSurface {
  template <typename Position, typename Temperature, typename Rate>
  SERAC_HOST_DEVICE auto operator()(double /*t*/, Position position, Temperature temperature, Rate dtemperature_dt) const
  {
    auto [X, dX_dxi] = position;
    auto [u, _0]     = temperature;
    auto [du_dt, _1] = dtemperature_dt;
    return X[0] + X[1] - cos(u) * du_dt;
  }
};

// ---- functional_boundary_test.cpp:103-111, 186-194 (the reference's lambdas as functors) ----------------------------------
template <int dim>
struct BoundaryTestH1 {
  double rho;
  template <typename Position, typename Temperature>
  SERAC_HOST_DEVICE auto operator()(double /*t*/, Position position, Temperature temperature) const
  {
    auto [X, dX_dxi] = position;
    auto [u, unused] = temperature;
    auto n           = normalize(cross(dX_dxi));
    tensor<double, dim> b{};
    b[0] = sin(X[0]);
    b[1] = X[0] * X[1];
    return X[0] * X[1] + dot(b, n) + rho * u;
  }
};
struct BoundaryTestL2 {
  double rho;
  template <typename Position, typename Temperature>
  SERAC_HOST_DEVICE auto operator()(double /*t*/, Position position, Temperature temperature) const
  {
    auto [X, dX_dxi] = position;
    auto [u, unused] = temperature;
    return X[0] * X[1] + rho * u;
  }
};
struct IdentityFunctor {  // bug_boundary_qoi.cpp:27-33
  template <typename Arg1, typename Arg2>
  SERAC_HOST_DEVICE auto operator()(Arg1, Arg2) const
  {
    return 1.0;
  }
};

// ---- functional_shape_derivatives.cpp:99-268: a vector field f(x) = C m(x), m = the monomials of degree <= p, C(i, j) =
// (i + 1) / (j + 1); the domain integrand is div f, the boundary integrand -f . n --------------------------------------------
template <int dim, int p>
struct MonomialField {
  static constexpr int nm = (dim == 2) ? ((p + 1) * (p + 2)) / 2 : ((p + 1) * (p + 2) * (p + 3)) / 6;
  /// exponents of monomial j in the reference's order: 1, x, y, (z,) x^2, xy, (xz,) y^2, (yz, z^2)
  static SERAC_HOST_DEVICE void exponents(int j, int (&e)[3])
  {
    const int e2[6][3]  = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {2, 0, 0}, {1, 1, 0}, {0, 2, 0}};
    const int e3[10][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {2, 0, 0}, {1, 1, 0}, {1, 0, 1}, {0, 2, 0}, {0, 1, 1}, {0, 0, 2}};
    for (int d = 0; d < 3; d++) e[d] = (dim == 2) ? e2[j][d] : e3[j][d];
  }
  template <typename T>
  static SERAC_HOST_DEVICE T power(T x, int n)
  {
    T r = x * 0.0 + 1.0;
    for (int i = 0; i < n; i++) r = r * x;
    return r;
  }
  template <typename T>
  static SERAC_HOST_DEVICE tensor<T, dim> value(const tensor<T, dim>& X)
  {
    tensor<T, dim> f{};
    for (int j = 0; j < nm; j++) {
      int e[3];
      exponents(j, e);
      T m = X[0] * 0.0 + 1.0;
      for (int d = 0; d < dim; d++) m = m * power(X[d], e[d]);
      for (int i = 0; i < dim; i++) f[i] = f[i] + (double(i + 1) / (j + 1)) * m;
    }
    return f;
  }
  template <typename T>
  static SERAC_HOST_DEVICE T divergence(const tensor<T, dim>& X)
  {
    T div = X[0] * 0.0;
    for (int j = 0; j < nm; j++) {
      int e[3];
      exponents(j, e);
      for (int i = 0; i < dim; i++) {  // d m_j / d x_i
        if (e[i] == 0) continue;
        T dm = X[0] * 0.0 + double(e[i]);
        for (int d = 0; d < dim; d++) dm = dm * power(X[d], d == i ? e[d] - 1 : e[d]);
        div = div + (double(i + 1) / (j + 1)) * dm;
      }
    }
    return div;
  }
};
template <int dim, int p>
struct DivergenceFunctor {  // TestFunctorOne, :245-252
  template <typename Position>
  SERAC_HOST_DEVICE auto operator()(double, Position position) const
  {
    return serac::tuple{MonomialField<dim, p>::divergence(get<VALUE>(position)), zero{}};
  }
};
template <int dim, int p>
struct NormalFluxFunctor {  // TestFunctorTwo, :254-264
  template <typename Position>
  SERAC_HOST_DEVICE auto operator()(double, Position position) const
  {
    auto [X, dX_dxi] = position;
    auto n           = normalize(cross(dX_dxi));
    return -dot(MonomialField<dim, p>::value(X), n);
  }
};

// a nonlinear integrand for the shape-aware check: source and flux depend on position, value and gradient
struct NonlinearThermal {
  template <typename Position, typename Temperature>
  SERAC_HOST_DEVICE auto operator()(double /*t*/, Position position, Temperature temperature) const
  {
    auto [X, dX_dxi] = position;
    auto [u, du_dX]  = temperature;
    auto source      = u * u + X[0] - 0.3 * X[2];
    auto flux        = (1.0 + 0.5 * u * u) * du_dX;
    return serac::tuple{source, flux};
  }
};

// ---- helpers ----------------------------------------------------------------------------------------------------------
mfem::Mesh make_mesh(mfem::Element::Type type, int n = 2)
{
  mfem::Mesh mesh = mfem::Mesh::MakeCartesian3D(n, n, n, type, 1.0, 1.2, 0.9);
  mesh.EnsureNodes();
  mesh.Transform([](const mfem::Vector& x, mfem::Vector& y) {
    y    = x;
    y(0) = x(0) + 0.05 * std::sin(2.0 * x(1)) * x(2);
    y(1) = x(1) + 0.04 * x(0) * x(2);
    y(2) = x(2) + 0.03 * std::cos(1.5 * x(0)) * x(1);
  });
  return mesh;
}

mfem::Mesh make_mesh_2d(mfem::Element::Type type, int n = 3)
{
  mfem::Mesh mesh = mfem::Mesh::MakeCartesian2D(n, n - 1, type, true, 1.0, 1.3);
  mesh.EnsureNodes();
  mesh.Transform([](const mfem::Vector& x, mfem::Vector& y) {
    y    = x;
    y(0) = x(0) + 0.05 * std::sin(2.0 * x(1));
    y(1) = x(1) + 0.04 * x(0) * x(0);
  });
  return mesh;
}

/// quadrilaterals below, triangles above (conforming), like the reference's patch2D_tris_and_quads.mesh
mfem::Mesh make_mixed_mesh_2d()
{
  mfem::Mesh mesh(2, 9, 6, 8);
  for (int j = 0; j < 3; j++)
    for (int i = 0; i < 3; i++) {
      const bool centre = (i == 1 && j == 1);
      mesh.AddVertex(0.6 * i + (centre ? 0.07 : 0.0) + 0.03 * j * i, 0.5 * j - (centre ? 0.04 : 0.0) + 0.02 * i * i);
    }
  mesh.AddQuad(0, 1, 4, 3);
  mesh.AddQuad(1, 2, 5, 4);
  mesh.AddTriangle(3, 4, 7);
  mesh.AddTriangle(3, 7, 6);
  mesh.AddTriangle(4, 5, 8);
  mesh.AddTriangle(4, 8, 7);
  const int bdr[8][2] = {{0, 1}, {1, 2}, {2, 5}, {5, 8}, {8, 7}, {7, 6}, {6, 3}, {3, 0}};
  for (int b = 0; b < 8; b++) mesh.AddBdrSegment(bdr[b][0], bdr[b][1], 1 + b / 2);
  mesh.FinalizeMesh();
  mesh.EnsureNodes();
  return mesh;
}

/// a patch of tetrahedra next to a patch of hexahedra (two pieces of one mesh, as in the reference's
/// patch3D_tets_and_hexes.mesh): triangular and quadrilateral boundary faces
mfem::Mesh make_mixed_mesh_3d(bool straight = false)
{
  mfem::Mesh tets  = mfem::Mesh::MakeCartesian3D(1, 1, 1, mfem::Element::TETRAHEDRON, 1.0, 1.1, 0.9);
  mfem::Mesh hexes = mfem::Mesh::MakeCartesian3D(2, 1, 1, mfem::Element::HEXAHEDRON, 1.0, 1.2, 0.8);
  if (straight) {
    hexes.Transform([](const mfem::Vector& x, mfem::Vector& y) {
      y    = x;
      y(0) = x(0) + 2.0;
    });
  } else {
    hexes.Transform([](const mfem::Vector& x, mfem::Vector& y) {
      y    = x;
      y(0) = x(0) + 2.0 + 0.05 * x(1) * x(2);
      y(2) = x(2) + 0.04 * x(0) * x(1);
    });
  }
  mfem::Mesh* pieces[] = {&tets, &hexes};
  mfem::Mesh  mesh(pieces, 2);
  mesh.EnsureNodes();
  return mesh;
}

double rel(const mfem::Vector& x, const mfem::Vector& y)
{
  mfem::Vector d(x);
  d -= y;
  return d.Norml2() / (y.Norml2() > 0 ? y.Norml2() : 1.0);
}

// ---- functional_comparisons.cpp ------------------------------------------------------------------------------------------
template <int p>
void thermal_comparison(mfem::Mesh& mesh, const char* tag)
{
  constexpr int dim = 3;
  mfem::H1_FECollection    fec(p, dim);
  mfem::FiniteElementSpace fes(&mesh, &fec);
  const mfem::FiniteElement&   el = *fes.GetFE(0);
  const mfem::IntegrationRule& ir = mfem::IntRules.Get(el.GetGeomType(), el.GetOrder() * 2);

  mfem::BilinearForm        A(&fes);
  mfem::ConstantCoefficient a_coef(a), b_coef(b);
  auto* mass = new mfem::MassIntegrator(a_coef);
  mass->SetIntRule(&ir);
  A.AddDomainIntegrator(mass);
  auto* diffusion = new mfem::DiffusionIntegrator(b_coef);
  diffusion->SetIntRule(&ir);
  A.AddDomainIntegrator(diffusion);
  A.Assemble(0);
  A.Finalize();
  mfem::LinearForm          f(&fes);
  mfem::FunctionCoefficient load_func([](const mfem::Vector& c) { return 100 * c(0) * c(1); });
  auto* load = new mfem::DomainLFIntegrator(load_func);
  load->SetIntRule(&ir);
  f.AddDomainIntegrator(load);
  f.Assemble();

  mfem::Vector U(fes.GetVSize());
  U.Randomize(1);

  using space = H1<p>;
  Functional<space(space)> residual(&fes, {&fes});
  residual.AddDomainIntegral(Dimension<dim>{}, DependsOn<0>{}, SourceFluxFunctor{}, mesh);

  mfem::Vector r1(U.Size());
  A.Mult(U, r1);
  r1 -= f;
  mfem::Vector r2 = residual(0.0, U);
  EXPECT(rel(r2, r1) < 1e-13, "thermal %s p=%d: ||r_functional - r_mfem|| / ||r|| = %.2e", tag, p, rel(r2, r1));

  auto [r3, drdU] = residual(0.0, differentiate_wrt(U));
  mfem::Vector dU(U.Size());
  dU.Randomize(2);
  mfem::Vector g1(U.Size());
  A.Mult(dU, g1);
  mfem::Vector g2 = drdU(dU);
  EXPECT(rel(g2, g1) < 1e-13, "thermal %s p=%d: gradient action vs A dU %.2e", tag, p, rel(g2, g1));
  // Gradient is an mfem::Operator
  const mfem::Operator& op = drdU;
  mfem::Vector          g3(op.Height());
  op.Mult(dU, g3);
  EXPECT(rel(g3, g1) < 1e-13 && op.Height() == fes.GetVSize() && op.Width() == fes.GetVSize(),
         "thermal %s p=%d: through mfem::Operator::Mult %.2e", tag, p, rel(g3, g1));
  std::unique_ptr<mfem::SparseMatrix> J = assemble(drdU);
  mfem::Vector g4(U.Size());
  J->Mult(dU, g4);
  EXPECT(rel(g4, g1) < 1e-13, "thermal %s p=%d: assembled matrix action %.2e", tag, p, rel(g4, g1));
  // entry by entry against mfem's matrix (same sparsity: both come from the element connectivity)
  const mfem::SparseMatrix& Am = A.SpMat();
  double                    worst = 0.0, scale = 0.0;
  for (int i = 0; i < Am.Height(); i++) {
    for (int k = Am.GetI()[i]; k < Am.GetI()[i + 1]; k++) {
      worst = std::max(worst, std::abs(Am.GetData()[k] - (*J)(i, Am.GetJ()[k])));
      scale = std::max(scale, std::abs(Am.GetData()[k]));
    }
  }
  EXPECT(worst < 1e-13 * scale && J->NumNonZeroElems() == Am.NumNonZeroElems(),
         "thermal %s p=%d: matrix entries max |diff| / max |A| = %.2e, nnz %d vs %d", tag, p, worst / scale,
         J->NumNonZeroElems(), Am.NumNonZeroElems());
}

// ---- functional_comparison_L2.cpp:56-160: L2 test and trial spaces against MFEM's mass matrix and load vector -----------
struct L2SourceFunctor {  // functional_comparison_L2.cpp:40-54 (a = 1.7, b = 0)
  template <typename Position, typename Temperature>
  SERAC_HOST_DEVICE auto operator()(double /*t*/, Position position, Temperature temperature) const
  {
    auto [X, dX_dxi] = position;
    auto [u, du_dX]  = temperature;
    auto source      = COEF_A * u - (100 * X[0] * X[1]);
    auto flux        = 0.0 * du_dX;
    return serac::tuple{source, flux};
  }
};

// the same comparison on a mesh of mixed element geometry: MFEM's integrators run element by element with the rule of
// each element's geometry (serac's rules: Q = p + 1 Gauss points per direction / the simplex rule of the same degree)
template <int p, int dim>
void thermal_comparison_mixed(mfem::Mesh& mesh, const char* tag)
{
  mfem::H1_FECollection    fec(p, dim);
  mfem::FiniteElementSpace fes(&mesh, &fec);
  const int                n = fes.GetVSize();
  mfem::SparseMatrix        A(n, n);
  mfem::Vector              f(n);
  f = 0.0;
  mfem::ConstantCoefficient a_coef(a), b_coef(b);
  mfem::FunctionCoefficient load_func([](const mfem::Vector& c) { return 100 * c(0) * c(1); });
  mfem::MassIntegrator      mass(a_coef);
  mfem::DiffusionIntegrator diffusion(b_coef);
  mfem::DomainLFIntegrator  load(load_func);
  mfem::Array<int>          dofs;
  for (int e = 0; e < mesh.GetNE(); e++) {
    const mfem::FiniteElement&   el = *fes.GetFE(e);
    mfem::ElementTransformation& T  = *mesh.GetElementTransformation(e);
    const mfem::IntegrationRule& ir = mfem::IntRules.Get(el.GetGeomType(), el.GetOrder() * 2);
    mass.SetIntRule(&ir);
    diffusion.SetIntRule(&ir);
    load.SetIntRule(&ir);
    mfem::DenseMatrix m, d;
    mfem::Vector      ev;
    mass.AssembleElementMatrix(el, T, m);
    diffusion.AssembleElementMatrix(el, T, d);
    m += d;
    load.AssembleRHSElementVect(el, T, ev);
    fes.GetElementDofs(e, dofs);
    A.AddSubMatrix(dofs, dofs, m);
    f.AddElementVector(dofs, ev);
  }
  A.Finalize();
  mfem::Vector U(n);
  U.Randomize(1);
  using space = H1<p>;
  Functional<space(space)> residual(&fes, {&fes});
  residual.AddDomainIntegral(Dimension<dim>{}, DependsOn<0>{}, SourceFluxFunctor{}, mesh);
  mfem::Vector r1(n);
  A.Mult(U, r1);
  r1 -= f;
  mfem::Vector r2 = residual(0.0, U);
  EXPECT(rel(r2, r1) < 1e-13, "thermal %s p=%d: ||r_functional - r_mfem|| / ||r|| = %.2e", tag, p, rel(r2, r1));
  auto [r3, drdU] = residual(0.0, differentiate_wrt(U));
  mfem::Vector dU(n);
  dU.Randomize(2);
  mfem::Vector g1(n);
  A.Mult(dU, g1);
  mfem::Vector g2 = drdU(dU);
  EXPECT(rel(g2, g1) < 1e-13, "thermal %s p=%d: gradient action vs A dU %.2e", tag, p, rel(g2, g1));
  std::unique_ptr<mfem::SparseMatrix> J = assemble(drdU);
  mfem::Vector                        g3(n);
  J->Mult(dU, g3);
  EXPECT(rel(g3, g1) < 1e-13, "thermal %s p=%d: assembled matrix action vs A dU %.2e", tag, p, rel(g3, g1));
  bool same_pattern = J->NumNonZeroElems() == A.NumNonZeroElems();
  for (int i = 0; same_pattern && i <= n; i++) same_pattern = J->GetI()[i] == A.GetI()[i];
  A.SortColumnIndices();
  for (int k = 0; same_pattern && k < A.NumNonZeroElems(); k++) same_pattern = J->GetJ()[k] == A.GetJ()[k];
  EXPECT(same_pattern, "thermal %s p=%d: sparsity pattern over both geometries equals MFEM's (%d nonzeros)", tag, p,
         J->NumNonZeroElems());
}

template <int p, int dim>
void l2_comparison(mfem::Mesh& mesh, const char* tag)
{
  // serac::generateParFiniteElementSpace<L2<p>>: L2_FECollection on Gauss-Lobatto points (finite_element.hpp / mesh utilities)
  mfem::L2_FECollection    fec(p, dim, mfem::BasisType::GaussLobatto);
  mfem::FiniteElementSpace fes(&mesh, &fec);
  const mfem::FiniteElement&   el = *fes.GetFE(0);
  const mfem::IntegrationRule& ir = mfem::IntRules.Get(el.GetGeomType(), 2 * p + 1);  // the (p + 1)-point Gauss rule

  mfem::BilinearForm        A(&fes);
  mfem::ConstantCoefficient a_coef(a);
  auto* mass = new mfem::MassIntegrator(a_coef);
  mass->SetIntRule(&ir);
  A.AddDomainIntegrator(mass);
  A.Assemble(0);
  A.Finalize();
  mfem::LinearForm          f(&fes);
  mfem::FunctionCoefficient load_func([](const mfem::Vector& c) { return 100 * c(0) * c(1); });
  auto* load = new mfem::DomainLFIntegrator(load_func);
  load->SetIntRule(&ir);
  f.AddDomainIntegrator(load);
  f.Assemble();

  mfem::Vector U(fes.GetVSize());
  U.Randomize(3);

  using space = L2<p>;
  Functional<space(space)> residual(&fes, {&fes});
  residual.AddDomainIntegral(Dimension<dim>{}, DependsOn<0>{}, L2SourceFunctor{}, mesh);

  mfem::Vector r1(U.Size());
  A.Mult(U, r1);
  r1 -= f;
  mfem::Vector r2 = residual(0.0, U);
  EXPECT(rel(r2, r1) < 1e-13, "L2 %s p=%d: ||r_functional - r_mfem|| / ||r|| = %.2e", tag, p, rel(r2, r1));

  auto [r3, drdU] = residual(0.0, differentiate_wrt(U));
  mfem::Vector g1(U.Size());
  A.Mult(U, g1);
  mfem::Vector g2 = drdU(U);
  EXPECT(rel(g2, g1) < 1e-13, "L2 %s p=%d: gradient action vs A U %.2e", tag, p, rel(g2, g1));
  std::unique_ptr<mfem::SparseMatrix> J = assemble(drdU);
  mfem::Vector g4(U.Size());
  J->Mult(U, g4);
  EXPECT(rel(g4, g1) < 1e-13 && J->NumNonZeroElems() == A.SpMat().NumNonZeroElems(),
         "L2 %s p=%d: assembled (block-diagonal) matrix action %.2e, nnz %d vs %d", tag, p, rel(g4, g1),
         J->NumNonZeroElems(), A.SpMat().NumNonZeroElems());
}

template <int p>
void elasticity_comparison(mfem::Mesh& mesh, const char* tag)
{
  constexpr int dim = 3;
  mfem::H1_FECollection    fec(p, dim);
  mfem::FiniteElementSpace fes(&mesh, &fec, dim, mfem::Ordering::byVDIM);
  const mfem::FiniteElement&   el = *fes.GetFE(0);
  const mfem::IntegrationRule& ir = mfem::IntRules.Get(el.GetGeomType(), el.GetOrder() * 2);

  mfem::BilinearForm        A(&fes);
  mfem::ConstantCoefficient a_coef(a), lambda_coef(b), mu_coef(b);
  auto* mass = new mfem::VectorMassIntegrator(a_coef);
  mass->SetIntRule(&ir);
  A.AddDomainIntegrator(mass);
  auto* elast = new mfem::ElasticityIntegrator(lambda_coef, mu_coef);
  elast->SetIntRule(&ir);
  A.AddDomainIntegrator(elast);
  A.Assemble(0);
  A.Finalize();
  mfem::LinearForm               f(&fes);
  mfem::VectorFunctionCoefficient load_func(dim, [](const mfem::Vector&, mfem::Vector& force) {
    force    = 0.0;
    force(0) = -1.0;
  });
  auto* load = new mfem::VectorDomainLFIntegrator(load_func);
  load->SetIntRule(&ir);
  f.AddDomainIntegrator(load);
  f.Assemble();

  mfem::Vector U(fes.GetVSize());
  U.Randomize(3);
  using space = H1<p, dim>;
  Functional<space(space)> residual(&fes, {&fes});
  residual.AddVolumeIntegral(DependsOn<0>{}, StressFunctor<dim>{}, mesh);
  mfem::Vector r1(U.Size());
  A.Mult(U, r1);
  r1 -= f;
  mfem::Vector r2 = residual(0.0, U);
  EXPECT(rel(r2, r1) < 1e-13, "elasticity %s p=%d: ||r_functional - r_mfem|| / ||r|| = %.2e", tag, p, rel(r2, r1));
  auto [r3, drdU] = residual(0.0, differentiate_wrt(U));
  mfem::Vector dU(U.Size());
  dU.Randomize(4);
  mfem::Vector g1(U.Size());
  A.Mult(dU, g1);
  mfem::Vector g2 = drdU(dU);
  EXPECT(rel(g2, g1) < 1e-13, "elasticity %s p=%d: gradient action vs A dU %.2e", tag, p, rel(g2, g1));
  std::unique_ptr<mfem::SparseMatrix> J = assemble(drdU);
  mfem::Vector g4(U.Size());
  J->Mult(dU, g4);
  EXPECT(rel(g4, g1) < 1e-13, "elasticity %s p=%d: assembled matrix action %.2e", tag, p, rel(g4, g1));
}

// ---- check_gradient.hpp:16-90 --------------------------------------------------------------------------------------------
template <typename F>
void check_gradient(F& f, double t, mfem::Vector& U, const char* what, double epsilon = 1.0e-4)
{
  mfem::Vector dU(U.Size());
  dU.Randomize(42);
  auto [value, dfdU]                              = f(t, differentiate_wrt(U));
  std::unique_ptr<mfem::SparseMatrix> dfdU_matrix = assemble(dfdU);
  mfem::Vector df_jvp1 = dfdU(dU);  // matrix-free
  mfem::Vector df_jvp2(df_jvp1.Size());
  dfdU_matrix->Mult(dU, df_jvp2);  // sparse matvec
  double relative_error = 0.0;
  if (df_jvp1.Norml2() != 0) relative_error = df_jvp1.DistanceTo(df_jvp2.GetData()) / df_jvp1.Norml2();
  EXPECT(relative_error < 5.e-6, "%s: JVP vs assembled matrix %.2e", what, relative_error);

  mfem::Vector f_values[5];
  for (int i = 0; i < 5; i++) {
    mfem::Vector U_plus_small = U;
    U_plus_small.Add((i - 2) * epsilon, dU);
    f_values[i] = f(t, U_plus_small);
  }
  mfem::Vector df1_fd[2], df1_cd[2];
  df1_fd[0] = f_values[4];
  df1_fd[0] -= f_values[2];
  df1_fd[0] /= (2.0 * epsilon);
  df1_fd[1] = f_values[3];
  df1_fd[1] -= f_values[2];
  df1_fd[1] /= epsilon;
  df1_cd[0] = f_values[4];
  df1_cd[0] -= f_values[0];
  df1_cd[0] /= (4.0 * epsilon);
  df1_cd[1] = f_values[3];
  df1_cd[1] -= f_values[1];
  df1_cd[1] /= (2.0 * epsilon);
  double denominator = (df_jvp1.Norml2() < 1.0e-10) ? 1.0 : df_jvp1.Norml2();
  double e1 = df1_fd[0].DistanceTo(df_jvp1.GetData()) / denominator;
  double e2 = df1_fd[1].DistanceTo(df_jvp1.GetData()) / denominator;
  EXPECT(std::fabs(e1 / e2 - 2.0) < 0.1 || std::fmin(e1, e2) < 1.0e-9, "%s: forward differences e1 %.2e e2 %.2e", what, e1, e2);
  double e3 = df1_cd[0].DistanceTo(df_jvp1.GetData()) / denominator;
  double e4 = df1_cd[1].DistanceTo(df_jvp1.GetData()) / denominator;
  EXPECT((std::fabs(e3 / e4 - 4.0) < 0.1) || std::fmin(e3, e4) < 1.0e-9, "%s: centred differences e3 %.2e e4 %.2e", what, e3, e4);
}

template <int p>
void elasticity_gradients(mfem::Mesh& mesh, const char* tag)  // functional_basic_h1_vector.cpp:119-141
{
  constexpr int dim = 3;
  using space       = H1<p, dim>;
  mfem::H1_FECollection    fec(p, dim);
  mfem::FiniteElementSpace fes(&mesh, &fec, dim, mfem::Ordering::byVDIM);
  mfem::Vector             U(fes.GetVSize());
  U.Randomize(5);
  Functional<space(space)> residual(&fes, {&fes});
  residual.AddDomainIntegral(Dimension<dim>{}, DependsOn<0>{}, ElasticityTestModelOne<dim>{}, mesh);
  residual.AddBoundaryIntegral(Dimension<dim - 1>{}, DependsOn<0>{}, ElasticityTestModelTwo<dim>{}, mesh);
  std::string what = std::string("elasticity_test ") + tag + " p=" + std::to_string(p);
  check_gradient(residual, 0.0, U, what.c_str());
}

template <int ptest, int ptrial, int dim>
void thermal_gradients(mfem::Mesh& mesh, const char* tag)  // functional_basic_h1_scalar.cpp:62-91
{
  using test_space  = H1<ptest>;
  using trial_space = H1<ptrial>;
  mfem::H1_FECollection    test_fec(ptest, dim), trial_fec(ptrial, dim);
  mfem::FiniteElementSpace test_fes(&mesh, &test_fec), trial_fes(&mesh, &trial_fec);
  mfem::GridFunction        U_gf(&trial_fes);
  mfem::FunctionCoefficient x_squared([](const mfem::Vector& x) { return x[0] * x[0]; });
  U_gf.ProjectCoefficient(x_squared);
  mfem::Vector U(U_gf);
  Functional<test_space(trial_space)> residual(&test_fes, {&trial_fes});
  residual.AddDomainIntegral(Dimension<dim>{}, DependsOn<0>{}, TestThermalModelOne<dim>{}, mesh);
  residual.AddBoundaryIntegral(Dimension<dim - 1>{}, DependsOn<0>{}, TestThermalModelTwo{}, mesh);
  std::string what = std::string("thermal_test ") + tag + " <" + std::to_string(ptest) + "," + std::to_string(ptrial) + ">";
  check_gradient(residual, 0.0, U, what.c_str());
}

// ---- functional_nonlinear.cpp:45-137 ---------------------------------------------------------------------------------------
template <int p, int dim>
void nonlinear_thermal(mfem::Mesh& mesh, const char* tag)
{
  using space = H1<p>;
  mfem::H1_FECollection    fec(p, dim);
  mfem::FiniteElementSpace fes(&mesh, &fec);
  mfem::Vector             U(fes.GetVSize());
  U.Randomize(11);
  Functional<space(space)> residual(&fes, {&fes});
  residual.AddDomainIntegral(Dimension<dim>{}, DependsOn<0>{}, NonlinearThermalDomain{}, mesh);
  residual.AddBoundaryIntegral(Dimension<dim - 1>{}, DependsOn<0>{}, NonlinearThermalBoundary{}, mesh);
  std::string what = std::string("nonlinear Thermal ") + tag + " p=" + std::to_string(p);
  check_gradient(residual, 0.0, U, what.c_str());
}
template <int p, int dim>
void nonlinear_elasticity(mfem::Mesh& mesh, const char* tag)
{
  using space = H1<p, dim>;
  mfem::H1_FECollection    fec(p, dim);
  mfem::FiniteElementSpace fes(&mesh, &fec, dim, mfem::Ordering::byVDIM);
  mfem::Vector             U(fes.GetVSize());
  U.Randomize(12);
  Functional<space(space)> residual(&fes, {&fes});
  residual.AddDomainIntegral(Dimension<dim>{}, DependsOn<0>{}, NonlinearElasticDomain{}, mesh);
  residual.AddBoundaryIntegral(Dimension<dim - 1>{}, DependsOn<0>{}, NonlinearElasticBoundary{}, mesh);
  std::string what = std::string("nonlinear Elasticity ") + tag + " p=" + std::to_string(p);
  check_gradient(residual, 0.0, U, what.c_str());
}

// ---- functional_multiphysics.cpp:25-83 + check_gradient.hpp:92-222 (two arguments) ----------------------------------------
template <typename F>
void check_gradient2(F& f, double t, const mfem::Vector& U, const mfem::Vector& dU_dt, const char* what, double epsilon = 1.0e-4)
{
  mfem::Vector dU(U.Size()), ddU_dt(U.Size());
  dU.Randomize(42);
  ddU_dt.Randomize(43);
  auto differences = [&](auto&& eval, const mfem::Vector& jvp, const char* which) {
    mfem::Vector f_values[5];
    for (int i = 0; i < 5; i++) f_values[i] = eval((i - 2) * epsilon);
    mfem::Vector fd0(f_values[4]), fd1(f_values[3]), cd0(f_values[4]), cd1(f_values[3]);
    fd0 -= f_values[2];
    fd0 /= (2.0 * epsilon);
    fd1 -= f_values[2];
    fd1 /= epsilon;
    cd0 -= f_values[0];
    cd0 /= (4.0 * epsilon);
    cd1 -= f_values[1];
    cd1 /= (2.0 * epsilon);
    const double den = (jvp.Norml2() < 1.0e-10) ? 1.0 : jvp.Norml2();
    const double e1 = fd0.DistanceTo(jvp.GetData()) / den, e2 = fd1.DistanceTo(jvp.GetData()) / den;
    const double e3 = cd0.DistanceTo(jvp.GetData()) / den, e4 = cd1.DistanceTo(jvp.GetData()) / den;
    EXPECT((std::fabs(e1 / e2 - 2.0) < 0.1 || std::fmin(e1, e2) < 1.0e-9) && (std::fabs(e3 / e4 - 4.0) < 0.1 || std::fmin(e3, e4) < 1.0e-9),
           "%s %s: forward e1 %.2e e2 %.2e, centred e3 %.2e e4 %.2e", what, which, e1, e2, e3, e4);
  };
  {
    auto [value, dfdU] = f(t, differentiate_wrt(U), dU_dt);
    std::unique_ptr<mfem::SparseMatrix> M = assemble(dfdU);
    mfem::Vector j1 = dfdU(dU), j2(j1.Size());
    M->Mult(dU, j2);
    const double re = j1.Norml2() != 0 ? j1.DistanceTo(j2.GetData()) / j1.Norml2() : 0.0;
    EXPECT(re < 5.e-6, "%s d/dU: JVP vs assembled matrix %.2e", what, re);
    differences([&](double h) { mfem::Vector Up(U); Up.Add(h, dU); return mfem::Vector(f(t, Up, dU_dt)); }, j1, "d/dU");
  }
  {
    auto [value, dfdV] = f(t, U, differentiate_wrt(dU_dt));
    std::unique_ptr<mfem::SparseMatrix> M = assemble(dfdV);
    mfem::Vector j1 = dfdV(ddU_dt), j2(j1.Size());
    M->Mult(ddU_dt, j2);
    const double re = j1.DistanceTo(j2.GetData()) / j1.Norml2();
    EXPECT(re < 5.e-14, "%s d/d(dU_dt): JVP vs assembled matrix %.2e", what, re);
    differences([&](double h) { mfem::Vector Vp(dU_dt); Vp.Add(h, ddU_dt); return mfem::Vector(f(t, U, Vp)); }, j1, "d/d(dU_dt)");
  }
}

void multiphysics(mfem::Mesh& mesh, const char* tag)
{
  constexpr int p = 3;
  using space     = H1<p>;
  mfem::H1_FECollection    fec(p, 3);
  mfem::FiniteElementSpace fes(&mesh, &fec);
  mfem::Vector             U(fes.GetVSize()), dU_dt(fes.GetVSize());
  U.Randomize(1);
  dU_dt.Randomize(2);
  Functional<space(space, space)> residual(&fes, {&fes, &fes});
  residual.AddVolumeIntegral(DependsOn<0, 1>{}, MultiphysicsVolume{}, mesh);
  residual.AddSurfaceIntegral(DependsOn<0, 1>{}, MultiphysicsSurface{}, mesh);
  std::string what = std::string("multiphysics NonlinearThermalTest3D ") + tag;
  check_gradient2(residual, 0.0, U, dU_dt, what.c_str());
}

// ---- functional_boundary_test.cpp:59-233 -------------------------------------------------------------------------------
template <int p, int dim, bool discontinuous>
void boundary_test(mfem::Mesh& mesh, const char* tag)
{
  const double rho = 1.75;
  using space      = std::conditional_t<discontinuous, L2<p>, H1<p>>;
  std::unique_ptr<mfem::FiniteElementCollection> fec;
  if (discontinuous)
    fec = std::make_unique<mfem::L2_FECollection>(p, dim, mfem::BasisType::GaussLobatto);
  else
    fec = std::make_unique<mfem::H1_FECollection>(p, dim);
  mfem::FiniteElementSpace fes(&mesh, fec.get());
  mfem::LinearForm                f(&fes);
  mfem::FunctionCoefficient       scalar_function([&](const mfem::Vector& coords) { return coords(0) * coords(1); });
  mfem::VectorFunctionCoefficient vector_function(dim, [&](const mfem::Vector& coords, mfem::Vector& output) {
    output    = 0.0;
    output[0] = sin(coords[0]);
    output[1] = coords[0] * coords[1];
  });
  mfem::BilinearForm        B(&fes);
  mfem::ConstantCoefficient density(rho);
  if (discontinuous) {
    f.AddBdrFaceIntegrator(new mfem::BoundaryLFIntegrator(scalar_function, 2, 0));
    B.AddBdrFaceIntegrator(new mfem::BoundaryMassIntegrator(density));
  } else {
    f.AddBoundaryIntegrator(new mfem::BoundaryLFIntegrator(scalar_function, 2, 0));
    f.AddBoundaryIntegrator(new mfem::BoundaryNormalLFIntegrator(vector_function, 2, 0));
    B.AddBoundaryIntegrator(new mfem::BoundaryMassIntegrator(density));
  }
  f.Assemble();
  B.Assemble(0);
  B.Finalize();
  mfem::Vector U(fes.GetVSize());
  if (discontinuous) {
    mfem::GridFunction        u_global(&fes);
    mfem::FunctionCoefficient xfunc([](const mfem::Vector& x) { return x[0]; });
    u_global.ProjectCoefficient(xfunc);
    U = u_global;
  } else {
    U.Randomize(3);
  }
  Functional<space(space)> residual(&fes, {&fes});
  if constexpr (discontinuous) {
    residual.AddBoundaryIntegral(Dimension<dim - 1>{}, DependsOn<0>{}, BoundaryTestL2{rho}, mesh);
  } else {
    residual.AddBoundaryIntegral(Dimension<dim - 1>{}, DependsOn<0>{}, BoundaryTestH1<dim>{rho}, mesh);
  }
  mfem::Vector r1(U.Size());
  B.Mult(U, r1);
  r1 += f;
  mfem::Vector r2 = residual(0.0, U);
  std::string what = std::string(discontinuous ? "boundaryL2 " : "FunctionalBoundary ") + tag + " p=" + std::to_string(p);
  if constexpr (!discontinuous) check_gradient(residual, 0.0, U, what.c_str());
  EXPECT(rel(r2, r1) < 1e-12, "%s: boundary residual vs MFEM's boundary integrators %.2e", what.c_str(), rel(r2, r1));
}

// ---- bug_boundary_qoi.cpp:37-64: ShapeAwareFunctional<shape, double()> on a boundary and on the faces of one attribute ------
void boundary_qoi_by_attribute()
{
  constexpr int ORDER = 1;
  mfem::Mesh    mesh  = mfem::Mesh::MakeCartesian2D(10, 10, mfem::Element::QUADRILATERAL, false, 1.0, 1.0);
  mesh.EnsureNodes();
  using shapeFES = H1<ORDER, 2>;
  mfem::H1_FECollection    shape_fec(ORDER, 2);
  mfem::FiniteElementSpace shape_fes(&mesh, &shape_fec, 2, mfem::Ordering::byVDIM);
  mfem::Vector             shape(shape_fes.GetVSize());
  shape = 0.0;
  ShapeAwareFunctional<shapeFES, double()> totalSurfArea(&shape_fes, {});
  totalSurfArea.AddBoundaryIntegral(Dimension<2 - 1>{}, DependsOn<>{}, IdentityFunctor{}, mesh);
  double totalSurfaceArea = totalSurfArea(0.0, shape);
  Domain attr1 = Domain::ofBoundaryElements(mesh, by_attr<2>(1));
  ShapeAwareFunctional<shapeFES, double()> attr1SurfArea(&shape_fes, {});
  attr1SurfArea.AddBoundaryIntegral(Dimension<2 - 1>{}, DependsOn<>{}, IdentityFunctor{}, attr1);
  double attr1SurfaceArea = attr1SurfArea(0.0, shape);
  EXPECT(std::fabs(totalSurfaceArea - 4.0) < 1.0e-14 && std::fabs(attr1SurfaceArea - 1.0) < 1.0e-14,
         "BoundaryIntegralQOI AttrBug: perimeter %.15g (4), boundary attribute 1 %.15g (1)", totalSurfaceArea, attr1SurfaceArea);
}

// ---- functional_shape_derivatives.cpp:266-345: the divergence theorem on the shape-displaced mesh --------------------------
template <int p, int dim>
void shape_derivative(mfem::Mesh& mesh, const char* tag, double tolerance)
{
  using test_space  = H1<p>;
  using trial_space = H1<p>;
  using shape_space = H1<p, dim>;
  mfem::H1_FECollection    fec(p, dim);
  mfem::FiniteElementSpace fes1(&mesh, &fec), fes2(&mesh, &fec, dim, mfem::Ordering::byVDIM);
  mfem::Vector ones(fes1.GetVSize()), U1(fes1.GetVSize()), U2(fes2.GetVSize()), dU2(fes2.GetVSize());
  ones = 1;
  U1.Randomize(1);
  U2.Randomize(2);
  U2 *= 0.1;
  dU2.Randomize(3);
  ShapeAwareFunctional<shape_space, test_space(trial_space)> residual(&fes2, &fes1, {&fes1});
  residual.AddDomainIntegral(Dimension<dim>{}, DependsOn<>{}, DivergenceFunctor<dim, p>{}, mesh);
  residual.AddBoundaryIntegral(Dimension<dim - 1>{}, DependsOn<>{}, NormalFluxFunctor<dim, p>{}, mesh);
  auto [r, drdU2] = residual(0.0, differentiate_wrt(U2), U1);
  const double sum_r = r * ones;
  mfem::Vector dr    = drdU2(dU2);
  const double sum_dr = ones * dr;
  EXPECT(std::fabs(sum_r) < tolerance && std::fabs(sum_dr) < tolerance,
         "ShapeDerivative %s p=%d: sum of the residual %.2e, of its shape derivative %.2e (tolerance %.1e)", tag, p, sum_r, sum_dr,
         tolerance);
}

// ---- functional_with_domain.cpp --------------------------------------------------------------------------------------------
template <int ptest, int ptrial, int dim>
void whole_mesh_comparison(mfem::Mesh& mesh, const char* tag)  // :89-146: integrals on complementary domains add up
{
  using test_space  = H1<ptest>;
  using trial_space = H1<ptrial>;
  mfem::H1_FECollection    test_fec(ptest, dim), trial_fec(ptrial, dim);
  mfem::FiniteElementSpace test_fes(&mesh, &test_fec), trial_fes(&mesh, &trial_fec);
  mfem::GridFunction        U_gf(&trial_fes);
  mfem::FunctionCoefficient x_squared([](const mfem::Vector& x) { return x[0] * x[0]; });
  U_gf.ProjectCoefficient(x_squared);
  mfem::Vector U(U_gf);
  Functional<test_space(trial_space)> residual(&test_fes, {&trial_fes});
  Functional<test_space(trial_space)> residual_comparison(&test_fes, {&trial_fes});
  // (the reference's meshes span [0,1]^d; the generated ones are split at their own mid-planes)
  mfem::Vector lo, hi;
  mesh.GetBoundingBox(lo, hi);
  const double xm = 0.5 * (lo(0) + hi(0)), ym = 0.5 * (lo(1) + hi(1));
  auto everything = [](std::vector<tensor<double, dim>>, int /* attr */) { return true; };
  auto on_left    = [xm](std::vector<tensor<double, dim>> X, int /* attr */) { return average(X)[0] < xm; };
  auto on_right   = [xm](std::vector<tensor<double, dim>> X, int /* attr */) { return average(X)[0] >= xm; };
  auto on_bottom  = [ym](std::vector<tensor<double, dim>> X, int /* attr */) { return average(X)[1] < ym; };
  auto on_top     = [ym](std::vector<tensor<double, dim>> X, int /* attr */) { return average(X)[1] >= ym; };
  Domain whole_mesh      = Domain::ofElements(mesh, everything);
  Domain left            = Domain::ofElements(mesh, on_left);
  Domain right           = Domain::ofElements(mesh, on_right);
  Domain whole_boundary  = Domain::ofBoundaryElements(mesh, everything);
  Domain bottom_boundary = Domain::ofBoundaryElements(mesh, on_bottom);
  Domain top_boundary    = Domain::ofBoundaryElements(mesh, on_top);
  residual.AddDomainIntegral(Dimension<dim>{}, DependsOn<0>{}, TestThermalIntegratorOne<dim>{}, left);
  residual.AddDomainIntegral(Dimension<dim>{}, DependsOn<0>{}, TestThermalIntegratorOne<dim>{}, right);
  residual.AddBoundaryIntegral(Dimension<dim - 1>{}, DependsOn<0>{}, TestThermalIntegratorTwo{}, bottom_boundary);
  residual.AddBoundaryIntegral(Dimension<dim - 1>{}, DependsOn<0>{}, TestThermalIntegratorTwo{}, top_boundary);
  std::string what = std::string("whole_mesh_comparison ") + tag + " <" + std::to_string(ptest) + "," + std::to_string(ptrial) + ">";
  check_gradient(residual, 0.0, U, what.c_str());
  mfem::Vector r0 = residual(0.0, U);
  residual_comparison.AddDomainIntegral(Dimension<dim>{}, DependsOn<0>{}, TestThermalIntegratorOne<dim>{}, whole_mesh);
  residual_comparison.AddBoundaryIntegral(Dimension<dim - 1>{}, DependsOn<0>{}, TestThermalIntegratorTwo{}, whole_boundary);
  mfem::Vector r1 = residual_comparison(0.0, U);
  const bool split = !left.elements_.empty() && !right.elements_.empty() && !bottom_boundary.elements_.empty() &&
                     !top_boundary.elements_.empty();
  EXPECT(split && r0.DistanceTo(r1.GetData()) < 1.0e-14, "%s: left + right + bottom + top vs whole mesh %.2e", what.c_str(),
         r0.DistanceTo(r1.GetData()));
}

template <int ptest, int ptrial, int dim>
void partial_mesh_comparison(mfem::Mesh& mesh, const char* tag)  // :181-226: a domain vs a masked integrand on the whole mesh
{
  using test_space  = H1<ptest>;
  using trial_space = H1<ptrial>;
  mfem::H1_FECollection    test_fec(ptest, dim), trial_fec(ptrial, dim);
  mfem::FiniteElementSpace test_fes(&mesh, &test_fec), trial_fes(&mesh, &trial_fec);
  mfem::GridFunction        U_gf(&trial_fes);
  mfem::FunctionCoefficient x_squared([](const mfem::Vector& x) { return x[0] * x[0]; });
  U_gf.ProjectCoefficient(x_squared);
  mfem::Vector U(U_gf);
  Functional<test_space(trial_space)> residual(&test_fes, {&trial_fes});
  Functional<test_space(trial_space)> residual_comparison(&test_fes, {&trial_fes});
  auto   on_left = [](std::vector<tensor<double, dim>> X, int /* attr */) { return average(X)[0] < 4.0; };
  Domain left    = Domain::ofElements(mesh, on_left);
  auto   on_top       = [](std::vector<tensor<double, dim>> X, int /* attr */) { return average(X)[1] >= 0.99; };
  Domain top_boundary = Domain::ofBoundaryElements(mesh, on_top);
  residual.AddDomainIntegral(Dimension<dim>{}, DependsOn<0>{}, TestThermalIntegratorOne<dim>{}, left);
  residual.AddBoundaryIntegral(Dimension<dim - 1>{}, DependsOn<0>{}, TestThermalIntegratorTwo{}, top_boundary);
  std::string what = std::string("partial_mesh_comparison ") + tag;
  check_gradient(residual, 0.0, U, what.c_str());
  mfem::Vector r0 = residual(0.0, U);
  residual_comparison.AddDomainIntegral(Dimension<dim>{}, DependsOn<0>{}, TestThermalIntegratorOne<dim, true>{}, mesh);
  residual_comparison.AddBoundaryIntegral(Dimension<dim - 1>{}, DependsOn<0>{}, TestThermalIntegratorTwo<true>{}, mesh);
  mfem::Vector r1 = residual_comparison(0.0, U);
  EXPECT(!left.elements_.empty() && int(left.elements_.size()) < mesh.GetNE() && !top_boundary.elements_.empty() &&
             r0.DistanceTo(r1.GetData()) < 1.0e-14,
         "%s: %zu of %d elements, %zu boundary faces: domain vs masked integrand %.2e", what.c_str(), left.elements_.size(),
         mesh.GetNE(), top_boundary.elements_.size(), r0.DistanceTo(r1.GetData()));
}

/// a beam [0,8] x [0,1] (x [0,1]) like the reference's beam-*.mesh files (straight sides: the masks cut at x = 4, y = 1)
mfem::Mesh make_beam(mfem::Element::Type type)
{
  const bool three_d = (type == mfem::Element::HEXAHEDRON || type == mfem::Element::TETRAHEDRON);
  mfem::Mesh mesh    = three_d ? mfem::Mesh::MakeCartesian3D(8, 1, 1, type, 8.0, 1.0, 1.0)
                               : mfem::Mesh::MakeCartesian2D(8, 1, type, true, 8.0, 1.0);
  mesh.EnsureNodes();
  return mesh;
}

void qoi_partial_domains()  // functional_with_domain.cpp:251-316
{
  constexpr int dim = 3;
  mfem::Mesh    mesh = make_beam(mfem::Element::HEXAHEDRON);
  using trial_space = H1<1>;
  mfem::H1_FECollection    trial_fec(1, dim);
  mfem::FiniteElementSpace trial_fes(&mesh, &trial_fec);
  mfem::GridFunction        U_gf(&trial_fes);
  mfem::FunctionCoefficient x_squared([](const mfem::Vector& x) { return x[0] * x[0]; });
  U_gf.ProjectCoefficient(x_squared);
  mfem::Vector U(U_gf);
  {
    Functional<double(trial_space)> qoi({&trial_fes});
    auto   on_top       = [](std::vector<tensor<double, 3>> X, int /* attr */) { return average(X)[1] >= 0.99; };
    Domain top_boundary = Domain::ofBoundaryElements(mesh, on_top);
    qoi.AddBoundaryIntegral(Dimension<dim - 1>{}, DependsOn</*nothing*/>{}, TrivialIntegrator{}, top_boundary);
    double area = qoi(0.0, U);
    EXPECT(std::fabs(area - 8.0) < 1.0e-14, "qoi partial_boundary: area of the top of the beam %.16g (8)", area);
  }
  {
    Functional<double(trial_space)> qoi({&trial_fes});
    auto   on_left = [](std::vector<tensor<double, dim>> X, int /* attr */) { return average(X)[0] < 4.0; };
    Domain left    = Domain::ofElements(mesh, on_left);
    qoi.AddDomainIntegral(Dimension<dim>{}, DependsOn</*nothing*/>{}, TrivialIntegrator{}, left);
    double volume = qoi(0.0, U);
    EXPECT(std::fabs(volume - 4.0) < 1.0e-14, "qoi partial_domain: volume of the left half of the beam %.16g (4)", volume);
  }
}

// ---- functional_qoi.cpp ------------------------------------------------------------------------------------------------
// check_gradient.hpp:225-262 (one argument) and :264-330 (two arguments): central difference of the scalar against the
// assembled gradient vector and against the directional derivative
template <typename F>
void check_qoi_gradient(F& f, double t, const mfem::Vector& U, const char* what)
{
  mfem::Vector dU(U.Size());
  dU.Randomize(42);
  const double epsilon = 1.0e-8;
  auto [unused, dfdU]  = f(t, differentiate_wrt(U));
  std::unique_ptr<mfem::Vector> dfdU_vec = assemble(dfdU);
  mfem::Vector U_plus(U), U_minus(U);
  U_plus.Add(epsilon, dU);
  U_minus.Add(-epsilon, dU);
  double df1 = (f(t, U_plus) - f(t, U_minus)) / (2 * epsilon);
  double df2 = (*dfdU_vec) * dU;
  double df3 = dfdU(dU);
  EXPECT(std::fabs((df1 - df2) / df1) < 2.e-5 && std::fabs((df1 - df3) / df1) < 2.e-5,
         "%s: central difference %.10e, gradient vector . dU %.10e, directional derivative %.10e", what, df1, df2, df3);
}
template <typename F>
void check_qoi_gradient(F& f, double t, const mfem::Vector& U1, const mfem::Vector& U2, const char* what)
{
  // (the reference steps by 1e-8; on the 3D box the scalar is ~2e5 and its derivative w.r.t. the second argument ~1, so a
  // difference over 2e-8 carries a relative rounding noise of ~1e-3: 1e-5 keeps both noise and truncation below 1e-6)
  const double epsilon = 1.0e-5;
  {
    mfem::Vector dU(U1.Size());
    dU.Randomize(42);
    auto [unused, dfdU] = f(t, differentiate_wrt(U1), U2);
    std::unique_ptr<mfem::Vector> v = assemble(dfdU);
    mfem::Vector Up(U1), Um(U1);
    Up.Add(epsilon, dU);
    Um.Add(-epsilon, dU);
    double df1 = (f(t, Up, U2) - f(t, Um, U2)) / (2 * epsilon), df2 = (*v) * dU, df3 = dfdU(dU);
    EXPECT(std::fabs((df1 - df2) / df1) < 2.e-5 && std::fabs((df1 - df3) / df1) < 2.e-5,
           "%s d/dU1: central difference %.10e, gradient vector . dU %.10e, directional derivative %.10e", what, df1, df2, df3);
  }
  {
    mfem::Vector dU(U2.Size());
    dU.Randomize(43);
    auto [unused, dfdU] = f(t, U1, differentiate_wrt(U2));
    std::unique_ptr<mfem::Vector> v = assemble(dfdU);
    mfem::Vector Up(U2), Um(U2);
    Up.Add(epsilon, dU);
    Um.Add(-epsilon, dU);
    double df1 = (f(t, U1, Up) - f(t, U1, Um)) / (2 * epsilon), df2 = (*v) * dU, df3 = dfdU(dU);
    EXPECT(std::fabs((df1 - df2) / df1) < 2.e-5 && std::fabs((df1 - df3) / df1) < 2.e-5,
           "%s d/dU2: central difference %.10e, gradient vector . dU %.10e, directional derivative %.10e", what, df1, df2, df3);
  }
}

// functional_qoi.cpp:129-182: the same quantities through MFEM's linear forms
double lf_mfem(mfem::Mesh& mesh, bool with_boundary, bool x_moment)
{
  mfem::ConstantCoefficient one(1.0);
  mfem::H1_FECollection     fec(1, mesh.Dimension());
  mfem::FiniteElementSpace  fes(&mesh, &fec);
  mfem::LinearForm          lf(&fes);
  lf.AddDomainIntegrator(new mfem::DomainLFIntegrator(one));
  if (with_boundary) lf.AddBoundaryIntegrator(new mfem::BoundaryLFIntegrator(one));
  lf.Assemble();
  mfem::GridFunction        gf(&fes);
  mfem::FunctionCoefficient x_coordinate([](const mfem::Vector& x) { return x[0]; });
  if (x_moment)
    gf.ProjectCoefficient(x_coordinate);
  else
    gf.ProjectCoefficient(one);
  return lf(gf);
}

/// functional_qoi.cpp:196-290 on a box (2D: [0,1]^2, 3D: [0,4] x [0,2] x [0,2]) whose measures are known in closed form
template <int p, int dim>
void qoi_test(mfem::Mesh& mesh, const char* tag)
{
  using trial_space = H1<p>;
  mfem::H1_FECollection     fec(p, dim);
  mfem::FiniteElementSpace  fes(&mesh, &fec);
  mfem::GridFunction        U_gf(&fes), V_gf(&fes);
  mfem::FunctionCoefficient x_squared([](const mfem::Vector& x) { return x[0] * x[0]; });
  mfem::FunctionCoefficient x_coord([](const mfem::Vector& x) { return x[0]; });
  U_gf.ProjectCoefficient(x_squared);
  V_gf.ProjectCoefficient(x_coord);
  mfem::Vector U(U_gf), V(V_gf);
  const double t = 0.0;
  const double expected_measure[] = {1.0, 16.0}, expected_moment[] = {0.5, 32.0}, expected_sum[] = {5.0, 56.0};
  {
    Functional<double(trial_space)> measure({&fes});
    measure.AddDomainIntegral(Dimension<dim>{}, DependsOn<>{}, TrivialIntegrator{}, mesh);
    const double m = measure(t, U);
    EXPECT(std::fabs(m - expected_measure[dim - 2]) / expected_measure[dim - 2] < 1e-10 &&
               std::fabs(m - lf_mfem(mesh, false, false)) / m < 1e-10,
           "qoi Measure %s p=%d: %.14g (exact %g, mfem %.14g)", tag, p, m, expected_measure[dim - 2], lf_mfem(mesh, false, false));
  }
  {
    Functional<double(trial_space)> x_moment({&fes});
    x_moment.AddDomainIntegral(Dimension<dim>{}, DependsOn<>{}, ZeroIndexIntegrator{}, mesh);
    const double m = x_moment(t, U);
    Functional<double(trial_space)> x_moment_2({&fes});
    x_moment_2.AddDomainIntegral(Dimension<dim>{}, DependsOn<0>{}, GetZeroIntegrator{}, mesh);
    const double m2 = x_moment_2(t, V);
    EXPECT(std::fabs(m - expected_moment[dim - 2]) / expected_moment[dim - 2] < 1e-10 &&
               std::fabs(m - lf_mfem(mesh, false, true)) / m < 1e-10 && std::fabs(m2 - m) / m < 1e-10,
           "qoi Moment %s p=%d: from the position %.14g, from the field x %.14g (exact %g)", tag, p, m, m2, expected_moment[dim - 2]);
  }
  {
    Functional<double(trial_space)> sum_of_measures({&fes});
    sum_of_measures.AddDomainIntegral(Dimension<dim>{}, DependsOn<>{}, TrivialIntegrator{}, mesh);
    sum_of_measures.AddBoundaryIntegral(Dimension<dim - 1>{}, DependsOn<>{}, TrivialIntegrator{}, mesh);
    const double m = sum_of_measures(t, U);
    EXPECT(std::fabs(m - expected_sum[dim - 2]) / expected_sum[dim - 2] < 1e-10 &&
               std::fabs(m - lf_mfem(mesh, true, false)) / m < 1e-10,
           "qoi SumOfMeasures %s p=%d: %.14g (exact %g)", tag, p, m, expected_sum[dim - 2]);
  }
  {
    Functional<double(trial_space)> f({&fes});
    f.AddDomainIntegral(Dimension<dim>{}, DependsOn<0>{}, SineIntegrator{}, mesh);
    f.AddBoundaryIntegral(Dimension<dim - 1>{}, DependsOn<0>{}, CosineIntegrator{}, mesh);
    std::string what = std::string("qoi Nonlinear ") + tag + " p=" + std::to_string(p);
    check_qoi_gradient(f, t, U, what.c_str());
  }
  {  // functional_qoi.cpp:292-335: two arguments
    mfem::GridFunction        U2_gf(&fes);
    mfem::FunctionCoefficient y([](const mfem::Vector& x) { return x[1]; });
    U2_gf.ProjectCoefficient(y);
    mfem::Vector U2(U2_gf);
    Functional<double(trial_space, trial_space)> f({&fes, &fes});
    f.AddDomainIntegral(Dimension<dim>{}, DependsOn<0, 1>{}, FourArgSineIntegrator{}, mesh);
    f.AddBoundaryIntegral(Dimension<dim - 1>{}, DependsOn<0, 1>{}, FourArgCosineIntegrator{}, mesh);
    std::string what = std::string("qoi Variadic ") + tag + " p=" + std::to_string(p);
    check_qoi_gradient(f, t, U, U2, what.c_str());
  }
}

void qoi_misc()
{
  const double t = 0.0;
  {  // functional_qoi.cpp:337-368 DependsOnVectorValuedInput: integral of |u| with u = (1, 2, 2)/... -> 3 * volume
    constexpr int dim = 3;
    mfem::Mesh    mesh = mfem::Mesh::MakeCartesian3D(2, 2, 2, mfem::Element::HEXAHEDRON, 1.0, 1.0, 1.0);
    mesh.EnsureNodes();
    using trial_space = H1<1, dim>;
    mfem::H1_FECollection    fec(1, dim);
    mfem::FiniteElementSpace fes(&mesh, &fec, dim, mfem::Ordering::byVDIM);
    mfem::Vector             U(fes.GetVSize());
    for (int n = 0; n < fes.GetNDofs(); n++) {
      U(n * dim + 0) = 1.0;
      U(n * dim + 1) = 2.0;
      U(n * dim + 2) = 2.0;
    }
    Functional<double(trial_space)> f({&fes});
    f.AddVolumeIntegral(DependsOn<0>{}, GetNormZeroIntegrator{}, mesh);
    const double v = f(t, U);
    EXPECT(std::fabs(v - 3.0) < 1e-12, "qoi DependsOnVectorValuedInput: integral of |u| over the unit cube %.14g (3)", v);
  }
  {  // :370-423 AddAreaIntegral / AddVolumeIntegral
    mfem::Mesh quads = mfem::Mesh::MakeCartesian2D(3, 3, mfem::Element::QUADRILATERAL, true, 1.0, 1.0);
    quads.EnsureNodes();
    mfem::H1_FECollection    fec2(1, 2);
    mfem::FiniteElementSpace fes2(&quads, &fec2);
    mfem::Vector             U2(fes2.GetVSize());
    U2 = 0.0;
    Functional<double(H1<1>)> area({&fes2});
    area.AddAreaIntegral(DependsOn<>{}, TrivialIntegrator{}, quads);
    mfem::Mesh hexes = mfem::Mesh::MakeCartesian3D(2, 2, 2, mfem::Element::HEXAHEDRON, 1.0, 2.0, 1.5);
    hexes.EnsureNodes();
    mfem::H1_FECollection    fec3(1, 3);
    mfem::FiniteElementSpace fes3(&hexes, &fec3);
    mfem::Vector             U3(fes3.GetVSize());
    U3 = 0.0;
    Functional<double(H1<1>)> volume({&fes3});
    volume.AddVolumeIntegral(DependsOn<>{}, TrivialIntegrator{}, hexes);
    const double a = area(t, U2), v = volume(t, U3);
    EXPECT(std::fabs(a - 1.0) < 1e-13 && std::fabs(v - 3.0) < 1e-13, "qoi AddAreaIntegral %.14g (1), AddVolumeIntegral %.14g (3)", a, v);
  }
  {  // :425-452 UsingL2: an L2 and an H1 argument, volume + surface of a box
    constexpr int dim = 3;
    mfem::Mesh    mesh = mfem::Mesh::MakeCartesian3D(2, 1, 1, mfem::Element::HEXAHEDRON, 2.0, 1.0, 1.0);
    mesh.EnsureNodes();
    using trial_space_0 = L2<0>;
    using trial_space_1 = H1<1, dim>;
    mfem::L2_FECollection    fec0(0, dim, mfem::BasisType::GaussLobatto);
    mfem::H1_FECollection    fec1(1, dim);
    mfem::FiniteElementSpace fes0(&mesh, &fec0), fes1(&mesh, &fec1, dim, mfem::Ordering::byVDIM);
    mfem::Vector             U0(fes0.GetVSize()), U1(fes1.GetVSize());
    U0 = 0.0;
    U1 = 0.0;
    Functional<double(trial_space_0, trial_space_1)> f({&fes0, &fes1});
    f.AddVolumeIntegral(DependsOn<1>{}, TrivialVariadicIntegrator{}, mesh);
    const double v = f(t, U0, U1);
    EXPECT(std::fabs(v - 2.0) < 1e-13, "qoi UsingL2: Functional<double(L2<0>, H1<1,3>)> volume %.14g (2)", v);
  }
}

// ---- ShapeAwareFunctional ------------------------------------------------------------------------------------------------
template <int p>
void shape_aware(mfem::Element::Type type, const char* tag)
{
  constexpr int dim = 3;
  using space       = H1<p>;
  using shape_space = H1<1, dim>;
  mfem::Mesh               mesh = make_mesh(type);
  mfem::H1_FECollection    fec(p, dim), shape_fec(1, dim);
  mfem::FiniteElementSpace fes(&mesh, &fec), shape_fes(&mesh, &shape_fec, dim, mfem::Ordering::byVDIM);
  mfem::Vector             U(fes.GetVSize()), S(shape_fes.GetVSize());
  U.Randomize(6);
  S.Randomize(7);
  S *= 0.03;
  ShapeAwareFunctional<shape_space, space(space)> shaped(&shape_fes, &fes, {&fes});
  shaped.AddDomainIntegral(Dimension<dim>{}, DependsOn<0>{}, NonlinearThermal{}, mesh);
  mfem::Vector r_shaped = shaped(0.0, S, U);
  // the same integrand on the mesh whose (order-1, byVDIM) nodes were moved by the shape displacement
  mfem::Mesh moved = make_mesh(type);
  {
    mfem::GridFunction& nodes = *moved.GetNodes();
    const bool by_vdim = nodes.FESpace()->GetOrdering() == mfem::Ordering::byVDIM;
    const int  nn      = nodes.FESpace()->GetNDofs();
    for (int n = 0; n < nn; n++)
      for (int c = 0; c < dim; c++) nodes(by_vdim ? n * dim + c : c * nn + n) += S(n * dim + c);
  }
  mfem::FiniteElementSpace fes_moved(&moved, &fec);
  Functional<space(space)> plain(&fes_moved, {&fes_moved});
  plain.AddDomainIntegral(Dimension<dim>{}, DependsOn<0>{}, NonlinearThermal{}, moved);
  mfem::Vector r_plain = plain(0.0, U);
  EXPECT(rel(r_shaped, r_plain) < 1e-12, "shape-aware %s p=%d: residual on X + p(X) vs plain Functional on the moved mesh %.2e",
         tag, p, rel(r_shaped, r_plain));
  // derivative w.r.t. the field through the shape-aware wrapper: JVP vs central difference
  auto [r, drdU] = shaped(0.0, S, differentiate_wrt(U));
  mfem::Vector dU(U.Size());
  dU.Randomize(8);
  mfem::Vector jvp = drdU(dU);
  const double eps = 1e-6;
  mfem::Vector Up(U), Um(U);
  Up.Add(eps, dU);
  Um.Add(-eps, dU);
  mfem::Vector fd = shaped(0.0, S, Up);
  mfem::Vector fm = shaped(0.0, S, Um);
  fd -= fm;
  fd /= (2.0 * eps);
  EXPECT(rel(jvp, fd) < 1e-7, "shape-aware %s p=%d: d/dU JVP vs central difference %.2e", tag, p, rel(jvp, fd));
}

int main(int argc, char** argv)
{
  const std::string which = argc > 1 ? argv[1] : "all";
  mfem::Mesh        hexes = make_mesh(mfem::Element::HEXAHEDRON);
  mfem::Mesh        tets  = make_mesh(mfem::Element::TETRAHEDRON);
  if (which == "comparisons" || which == "all") {
    thermal_comparison<1>(hexes, "hexes");
    thermal_comparison<2>(hexes, "hexes");
    thermal_comparison<3>(hexes, "hexes");
    // (the reference compares against MFEM's integrators on tensor-product meshes only: on simplices of order >= 2 serac's
    // quadrature rules (polynomials.hpp:57-260) and mfem's IntRules differ; p = 1 integrands are exact under both)
    thermal_comparison<1>(tets, "tets");
    elasticity_comparison<1>(hexes, "hexes");
    elasticity_comparison<2>(hexes, "hexes");
    elasticity_comparison<1>(tets, "tets");
    // meshes of mixed element geometry against MFEM's integrators (which dispatch per element themselves)
    mfem::Mesh tris_and_quads = make_mixed_mesh_2d();
    mfem::Mesh tets_and_hexes = make_mixed_mesh_3d();
    thermal_comparison_mixed<1, 2>(tris_and_quads, "tris_and_quads");
    thermal_comparison_mixed<1, 3>(tets_and_hexes, "tets_and_hexes");
  }
  if (which == "l2" || which == "all") {  // functional_comparison_L2.cpp:162-170
    mfem::Mesh quads = mfem::Mesh::MakeCartesian2D(3, 2, mfem::Element::QUADRILATERAL, true, 1.0, 1.3);
    quads.EnsureNodes();
    l2_comparison<1, 2>(quads, "quads");
    l2_comparison<2, 2>(quads, "quads");
    l2_comparison<1, 3>(hexes, "hexes");
    l2_comparison<2, 3>(hexes, "hexes");
    l2_comparison<3, 3>(hexes, "hexes");
  }
  if (which == "gradients" || which == "all") {
    elasticity_gradients<1>(hexes, "hexes");
    elasticity_gradients<2>(hexes, "hexes");
    elasticity_gradients<1>(tets, "tets");
    elasticity_gradients<2>(tets, "tets");
  }
  if (which == "scalar" || which == "all") {  // functional_basic_h1_scalar.cpp:107-118 (2D: boundary integrals on segments)
    mfem::Mesh quads = make_mesh_2d(mfem::Element::QUADRILATERAL);
    mfem::Mesh tris  = make_mesh_2d(mfem::Element::TRIANGLE);
    thermal_gradients<1, 1, 3>(hexes, "hexes");
    thermal_gradients<1, 1, 2>(tris, "tris");
    thermal_gradients<1, 1, 3>(tets, "tets");
    thermal_gradients<1, 1, 2>(quads, "quads");
    thermal_gradients<2, 1, 2>(quads, "quads");
    thermal_gradients<2, 1, 2>(tris, "tris");
    thermal_gradients<2, 1, 3>(hexes, "hexes");
    thermal_gradients<2, 1, 3>(tets, "tets");
    thermal_gradients<2, 2, 2>(quads, "quads");
    // meshes of mixed element geometry (functional_basic_h1_scalar.cpp:112-117): one kernel and one restriction per geometry
    mfem::Mesh tris_and_quads = make_mixed_mesh_2d();
    mfem::Mesh tets_and_hexes = make_mixed_mesh_3d();
    thermal_gradients<1, 1, 2>(tris_and_quads, "tris_and_quads");
    thermal_gradients<1, 1, 3>(tets_and_hexes, "tets_and_hexes");
    thermal_gradients<2, 1, 2>(tris_and_quads, "tris_and_quads");
    thermal_gradients<2, 1, 3>(tets_and_hexes, "tets_and_hexes");
  }
  if (which == "domains" || which == "all") {  // functional_with_domain.cpp:160-249
    mfem::Mesh quads = make_mesh_2d(mfem::Element::QUADRILATERAL);
    mfem::Mesh tris  = make_mesh_2d(mfem::Element::TRIANGLE);
    mfem::Mesh tris_and_quads = make_mixed_mesh_2d();
    mfem::Mesh tets_and_hexes = make_mixed_mesh_3d();
    whole_mesh_comparison<1, 1, 2>(tris, "tris");
    whole_mesh_comparison<1, 1, 2>(quads, "quads");
    whole_mesh_comparison<1, 1, 2>(tris_and_quads, "tris_and_quads");
    whole_mesh_comparison<1, 1, 3>(tets, "tets");
    whole_mesh_comparison<1, 1, 3>(hexes, "hexes");
    whole_mesh_comparison<1, 1, 3>(tets_and_hexes, "tets_and_hexes");
    whole_mesh_comparison<2, 1, 2>(tris_and_quads, "tris_and_quads");
    whole_mesh_comparison<2, 1, 3>(tets_and_hexes, "tets_and_hexes");
    mfem::Mesh beam_tri = make_beam(mfem::Element::TRIANGLE), beam_quad = make_beam(mfem::Element::QUADRILATERAL);
    mfem::Mesh beam_tet = make_beam(mfem::Element::TETRAHEDRON), beam_hex = make_beam(mfem::Element::HEXAHEDRON);
    partial_mesh_comparison<1, 1, 2>(beam_tri, "tris");
    partial_mesh_comparison<1, 1, 2>(beam_quad, "quads");
    partial_mesh_comparison<1, 1, 3>(beam_tet, "tets");
    partial_mesh_comparison<1, 1, 3>(beam_hex, "hexes");
    qoi_partial_domains();
  }
  if (which == "qoi" || which == "all") {  // functional_qoi.cpp:581-604
    mfem::Mesh mesh2D = mfem::Mesh::MakeCartesian2D(4, 4, mfem::Element::QUADRILATERAL, true, 1.0, 1.0);
    mfem::Mesh mesh3D = mfem::Mesh::MakeCartesian3D(4, 2, 2, mfem::Element::HEXAHEDRON, 4.0, 2.0, 2.0);
    mfem::Mesh tris2D = mfem::Mesh::MakeCartesian2D(3, 3, mfem::Element::TRIANGLE, true, 1.0, 1.0);
    mfem::Mesh tets3D = mfem::Mesh::MakeCartesian3D(2, 1, 1, mfem::Element::TETRAHEDRON, 4.0, 2.0, 2.0);
    for (mfem::Mesh* m : {&mesh2D, &mesh3D, &tris2D, &tets3D}) m->EnsureNodes();
    qoi_test<1, 2>(mesh2D, "quads");
    qoi_test<2, 2>(mesh2D, "quads");
    qoi_test<1, 3>(mesh3D, "hexes");
    qoi_test<2, 3>(mesh3D, "hexes");
    qoi_test<1, 2>(tris2D, "tris");
    qoi_test<1, 3>(tets3D, "tets");
    qoi_misc();
  }
  if (which == "nonlinear" || which == "all") {  // functional_nonlinear.cpp:139-153, functional_multiphysics.cpp:25-83
    mfem::Mesh mesh2D = make_mixed_mesh_2d();
    mfem::Mesh mesh3D = make_mixed_mesh_3d();
    nonlinear_thermal<1, 2>(mesh2D, "2D tris_and_quads");
    nonlinear_thermal<2, 2>(mesh2D, "2D tris_and_quads");
    nonlinear_thermal<3, 2>(mesh2D, "2D tris_and_quads");
    nonlinear_thermal<1, 3>(mesh3D, "3D tets_and_hexes");
    nonlinear_thermal<2, 3>(mesh3D, "3D tets_and_hexes");
    nonlinear_thermal<3, 3>(mesh3D, "3D tets_and_hexes");
    nonlinear_elasticity<1, 2>(mesh2D, "2D tris_and_quads");
    nonlinear_elasticity<2, 2>(mesh2D, "2D tris_and_quads");
    nonlinear_elasticity<3, 2>(mesh2D, "2D tris_and_quads");
    nonlinear_elasticity<1, 3>(mesh3D, "3D tets_and_hexes");
    nonlinear_elasticity<2, 3>(mesh3D, "3D tets_and_hexes");
    nonlinear_elasticity<3, 3>(mesh3D, "3D tets_and_hexes");
    multiphysics(mesh3D, "tets_and_hexes p=3");
  }
  if (which == "boundary" || which == "all") {  // functional_boundary_test.cpp:223-232, bug_boundary_qoi.cpp
    mfem::Mesh mesh2D = make_mixed_mesh_2d();
    mfem::Mesh mesh3D = make_mixed_mesh_3d(true);
    boundary_test<1, 2, false>(mesh2D, "2D tris_and_quads");
    boundary_test<2, 2, false>(mesh2D, "2D tris_and_quads");
    boundary_test<1, 3, false>(mesh3D, "3D tets_and_hexes");
    boundary_test<2, 3, false>(mesh3D, "3D tets_and_hexes");
    boundary_test<1, 2, true>(mesh2D, "2D tris_and_quads");
    boundary_test<2, 2, true>(mesh2D, "2D tris_and_quads");
    boundary_test<1, 3, true>(mesh3D, "3D tets_and_hexes");
    boundary_test<2, 3, true>(mesh3D, "3D tets_and_hexes");
    boundary_test<3, 3, true>(mesh3D, "3D tets_and_hexes");
    boundary_qoi_by_attribute();
  }
  if (which == "shape" || which == "all") {
    {  // functional_shape_derivatives.cpp:347-352 (the reference's tolerances)
      mfem::Mesh mesh2D = make_mixed_mesh_2d();
      mfem::Mesh mesh3D = make_mixed_mesh_3d();
      shape_derivative<1, 2>(mesh2D, "2D tris_and_quads", 3.0e-14);
      shape_derivative<2, 2>(mesh2D, "2D tris_and_quads", 3.0e-14);
      shape_derivative<1, 3>(mesh3D, "3D tets_and_hexes", 2.0e-13);
      shape_derivative<2, 3>(mesh3D, "3D tets_and_hexes", 1.5e-2);
    }
    shape_aware<2>(mfem::Element::HEXAHEDRON, "hexes");
    shape_aware<1>(mfem::Element::TETRAHEDRON, "tets");
  }
  std::printf("%s: %d failure(s)\n", which.c_str(), g_failures);
  return g_failures == 0 ? 0 : 1;
}
