// Test driver for the C++ host side (serac_b200/functional.hpp): builds serac::Functional objects the way the
// reference's physics modules and unit tests do (src/serac/physics/solid_mechanics.hpp:879-894, :1075-1101;
// src/serac/numerics/functional/tests/functional_basic_h1_scalar.cpp, check_gradient.hpp) on problems exported by
// pytest from the golden fixtures, and writes its results back for pytest to compare with the reference's outputs.
//
//   functional_test <case> <directory>
//     solid     H1<2,3>(H1<1,3>, H1<2,3>, H1<2,3>), shape-aware NeoHookean stress integrand (hex or tet mesh)
//     j2        same signature, J2SmallStrain<LinearHardening> with QuadratureData state, updateQdata(true)
//     j2_finite finite-deformation J2 with a default-constructed state buffer; checkpoint / restart of the state
//     pressure  same signature, ONE boundary integral: follower pressure (hex mesh, quadrilateral faces)
//     user      H1<2>(H1<2>) with a q-function defined in THIS file: checked here against finite differences and the
//               assembled matrix (check_gradient.hpp:16-90)
//     newton    the solid problem solved the way SolidMechanics::quasiStaticSolve does: Newton on the residual with the
//               essential rows zeroed, Jacobian = assemble + EliminateRowsCols, CG + Jacobi -- every step on the device;
//               once with the assembled matrix and once matrix-free over Gradient::Mult
//     equation_solver  the reference's EquationSolverSuite (numerics/tests/equationsolver.cpp): Newton + CG / GMRES
//               + Jacobi on sin-reaction-diffusion, known answer
//     krylov    CGSolver / GMRESSolver (assembled and matrix-free) on the eliminated Jacobian at the fixture's state
//     qoi       Functional<double(H1<2,3>)> with a q-function defined in this file: value, directional derivative, gradient
// Directory layout: <name>.bin raw little-endian arrays (int32 / float64), meta.txt "key value" lines.
#include <cmath>
#include <cstdio>
#include <fstream>
#include <map>
#include <string>

#include "serac_b200/functional.hpp"
#include "serac_b200/solvers.hpp"

using namespace serac;

static std::string g_dir;

template <typename T>
std::vector<T> read_bin(const std::string& name)
{
  std::ifstream f(g_dir + "/" + name + ".bin", std::ios::binary | std::ios::ate);
  if (!f) {
    std::fprintf(stderr, "missing input %s\n", name.c_str());
    std::exit(2);
  }
  const std::streamsize n = f.tellg();
  f.seekg(0);
  std::vector<T> v(size_t(n) / sizeof(T));
  f.read(reinterpret_cast<char*>(v.data()), n);
  return v;
}

template <typename T>
void write_bin(const std::string& name, const std::vector<T>& v)
{
  std::ofstream f(g_dir + "/out_" + name + ".bin", std::ios::binary);
  f.write(reinterpret_cast<const char*>(v.data()), std::streamsize(v.size() * sizeof(T)));
}

std::map<std::string, double> read_meta()
{
  std::map<std::string, double> m;
  std::ifstream                 f(g_dir + "/meta.txt");
  std::string                   k;
  double                        v;
  while (f >> k >> v) m[k] = v;
  return m;
}

struct Problem {
  Mesh               mesh;
  FiniteElementSpace space, shape_space, scalar_space;
};

/// spaces of one exported fixture: H1<p,dim> (u, a), H1<1,dim> (shape, and the mesh nodes), H1<p> (scalar)
Problem load_problem(const std::map<std::string, double>& meta, bool with_faces)
{
  Problem        P;
  const Geometry g   = Geometry(int(meta.at("geom")));
  const int      p   = int(meta.at("p"));
  const int      dim = int(meta.at("dim"));
  const int64_t  ne  = int64_t(meta.at("ne"));
  P.space            = FiniteElementSpace{g, p, dim, Ordering::byVDIM, int64_t(meta.at("num_nodes")), ne,
                               read_bin<int>("native_dofs")};
  P.scalar_space      = P.space;
  P.scalar_space.vdim = 1;
  P.shape_space = FiniteElementSpace{g, 1, dim, Ordering::byVDIM, int64_t(meta.at("shape_num_nodes")), ne,
                                     read_bin<int>("shape_native_dofs")};
  P.mesh.nodes_space = FiniteElementSpace{g, int(meta.at("mesh_order")), dim, Ordering(int(meta.at("mesh_ordering"))),
                                          int64_t(meta.at("mesh_num_nodes")), ne, read_bin<int>("mesh_native_dofs")};
  P.mesh.nodes       = read_bin<double>("mesh_nodes");
  if (with_faces) {
    const int64_t nf = int64_t(meta.at("nf"));
    P.space.num_boundary_faces = P.shape_space.num_boundary_faces = P.mesh.nodes_space.num_boundary_faces = nf;
    P.space.native_boundary_face_dofs            = read_bin<int>("face_native_dofs");
    P.shape_space.native_boundary_face_dofs      = read_bin<int>("shape_face_native_dofs");
    P.mesh.nodes_space.native_boundary_face_dofs = read_bin<int>("mesh_face_native_dofs");
  }
  return P;
}

constexpr int p   = 2;
constexpr int dim = 3;
using solid_test  = H1<p, dim>;
using SolidFunctional = Functional<solid_test(H1<1, dim>, H1<p, dim>, H1<p, dim>)>;

template <typename Material>
using SolidIntegrand = ShapeAwareIntegrandWrapperWithState<MaterialStressFunctor<Material>, dim, solid_test, H1<p, dim>, H1<p, dim>>;

int run_solid(const std::map<std::string, double>& meta)
{
  Problem P = load_problem(meta, false);
  auto    params = read_bin<double>("params");
  SolidFunctional residual(&P.space, {&P.shape_space, &P.space, &P.space});
  using Material = solid_mechanics::NeoHookean;
  Material m{params[0], params[1], params[2]};
  auto     gn = params[3] != 0.0 ? GeometricNonlinearities::On : GeometricNonlinearities::Off;
  Domain   whole = EntireDomain(P.mesh);
  if (meta.count("domain_elements")) {  // integral on part of the mesh (Domain::ofElements, functional_with_domain.cpp)
    whole.entire   = false;
    whole.elements = read_bin<int>("domain_elements");
  }
  // SolidMechanics::setMaterial (solid_mechanics.hpp:879-894)
  residual.AddDomainIntegral(Dimension<dim>{}, DependsOn<0, 1, 2>{}, SolidIntegrand<Material>(MaterialStressFunctor<Material>(m, gn)),
                             whole, EmptyQData);
  Vector S(read_bin<double>("S")), U(read_bin<double>("U")), A(read_bin<double>("A")), dU(read_bin<double>("dU"));
  // plain residual, then residual + Jacobian (functional.hpp:446-513)
  Vector& r0 = residual(0.0, S, U, A);
  write_bin("r_plain", r0.to_host());
  auto [r, drdu] = residual(0.0, S, differentiate_wrt(U), A);
  write_bin("r", r.to_host());
  write_bin("dr", drdu(dU).to_host());
  auto K = assemble(drdu);
  write_bin("row_ptr", K->row_ptr);
  write_bin("col_ind", K->col_ind);
  write_bin("values", K->values);
  // mass matrix action: derivative w.r.t. the acceleration
  auto [r2, drda] = residual(0.0, S, U, differentiate_wrt(A));
  write_bin("dr_da", drda(dU).to_host());
  return 0;
}

/// One quasi-static solve (solid_mechanics.hpp:1120-1140 residual / Jacobian lambdas, Newton + CG/Jacobi of
/// equation_solver.cpp) without leaving the GPU.  Inputs beyond the solid case: ess.bin (essential dofs), meta
/// force_scale / rel_tol / max_iter / newton_rel / newton_max.  matrix_free: CG over Gradient::Mult instead of the matrix.
int run_newton(const std::map<std::string, double>& meta, bool matrix_free, bool use_gmres = false)
{
  Problem P = load_problem(meta, false);
  auto    params = read_bin<double>("params");
  SolidFunctional residual(&P.space, {&P.shape_space, &P.space, &P.space});
  using Material = solid_mechanics::NeoHookean;
  Material m{params[0], params[1], params[2]};
  auto     gn = params[3] != 0.0 ? GeometricNonlinearities::On : GeometricNonlinearities::Off;
  Domain   whole = EntireDomain(P.mesh);
  residual.AddDomainIntegral(Dimension<dim>{}, DependsOn<0, 1, 2>{}, SolidIntegrand<Material>(MaterialStressFunctor<Material>(m, gn)),
                             whole, EmptyQData);
  DofList ess(read_bin<int>("ess"));
  Vector  S(read_bin<double>("S")), A(read_bin<double>("A"));
  {
    auto a = A.to_host();
    for (auto& v : a) v *= meta.at("force_scale");
    A.set(a.data());
  }
  const int64_t n = P.space.GetVSize();
  Vector        u(n), du(n), rhs(n);
  u = 0.0;
  CGSolver cg;
  cg.SetRelTol(meta.at("rel_tol"));
  cg.SetAbsTol(0.0);
  cg.SetMaxIter(int(meta.at("max_iter")));
  cg.iterative_mode = false;  // equation_solver.cpp:96: the Newton driver's linear solver starts from zero
  GMRESSolver gmres;
  gmres.SetRelTol(meta.at("rel_tol"));
  gmres.SetAbsTol(0.0);
  gmres.SetMaxIter(int(meta.at("max_iter")));
  gmres.iterative_mode = false;
  std::vector<double> norms, cg_iters;
  for (int it = 0; it < int(meta.at("newton_max")); it++) {
    auto [r, drdu] = residual(0.0, S, differentiate_wrt(u), A);
    rhs            = r;
    SetSubVector(rhs, ess, 0.0);  // r.SetSubVector(bcs_.allEssentialTrueDofs(), 0.0)
    norms.push_back(Norml2(rhs));
    if (norms.back() < meta.at("newton_rel") * norms.front()) break;
    auto J   = assemble_on_device(drdu);  // J_ = assemble(drdu)
    auto J_e = J->EliminateRowsCols(ess);  // J_e_ = bcs_.eliminateAllEssentialDofsFromMatrix(*J_)
    JacobiSmoother   prec(*J);
    GradientOperator op(drdu, ess);
    bool converged = false;
    int  iters     = 0;
    if (use_gmres) {  // serac's default: GMRES (solid_mechanics.hpp:50-57)
      gmres.SetOperator(*J);
      gmres.SetPreconditioner(prec);
      gmres.Mult(rhs, du);
      converged = gmres.GetConverged();
      iters     = gmres.GetNumIterations();
    } else {
      if (matrix_free)
        cg.SetOperator(op);
      else
        cg.SetOperator(*J);
      cg.SetPreconditioner(prec);
      cg.Mult(rhs, du);
      converged = cg.GetConverged();
      iters     = cg.GetNumIterations();
    }
    if (!converged) {
      std::fprintf(stderr, "the linear solver did not converge in Newton step %d\n", it);
      return 1;
    }
    cg_iters.push_back(double(iters));
    Add(u, -1.0, du);  // u -= du
  }
  const std::string tag = use_gmres ? "_gmres" : (matrix_free ? "_mf" : "");
  write_bin("u" + tag, u.to_host());
  write_bin("residual_norms" + tag, norms);
  write_bin("cg_iterations" + tag, cg_iters);
  std::printf("newton%s: %zu residual evaluations, last |r| = %.3e, linear iterations", tag.c_str(), norms.size(), norms.back());
  for (double c : cg_iters) std::printf(" %d", int(c));
  std::printf("\n");
  return 0;
}

/// The linear solvers on the Jacobian at the fixture's state, for comparison with the reference's mfem on the same
/// system (tests/golden/krylov_*.npz): inputs beyond the solid case: ess.bin, B_bc.bin (right-hand side after
/// EliminateBC), X.bin (initial guess), meta rel_tol / max_iter / jacobi.
int run_krylov(const std::map<std::string, double>& meta)
{
  Problem P = load_problem(meta, false);
  auto    params = read_bin<double>("params");
  SolidFunctional residual(&P.space, {&P.shape_space, &P.space, &P.space});
  using Material = solid_mechanics::NeoHookean;
  Material m{params[0], params[1], params[2]};
  auto     gn    = params[3] != 0.0 ? GeometricNonlinearities::On : GeometricNonlinearities::Off;
  Domain   whole = EntireDomain(P.mesh);
  residual.AddDomainIntegral(Dimension<dim>{}, DependsOn<0, 1, 2>{}, SolidIntegrand<Material>(MaterialStressFunctor<Material>(m, gn)),
                             whole, EmptyQData);
  DofList ess(read_bin<int>("ess"));
  Vector  S(read_bin<double>("S")), U(read_bin<double>("U")), A(read_bin<double>("A")), B(read_bin<double>("B_bc")),
      X(read_bin<double>("X"));
  auto [r, drdu] = residual(0.0, S, differentiate_wrt(U), A);
  auto J   = assemble_on_device(drdu);
  auto J_e = J->EliminateRowsCols(ess);
  write_bin("values_elim", J->values().to_host());
  write_bin("values_e", J_e->values().to_host());
  JacobiSmoother   prec(*J);
  GradientOperator op(drdu, ess);
  const bool       jacobi = meta.at("jacobi") != 0.0;
  const int64_t    n      = P.space.GetVSize();
  Vector           x(n);
  auto report = [&](const std::string& name, auto& solver) {
    write_bin("x_" + name, x.to_host());
    write_bin("info_" + name, std::vector<double>{double(solver.GetConverged()), double(solver.GetNumIterations()),
                                                  solver.GetInitialNorm(), solver.GetFinalNorm()});
    std::printf("%s: converged %d, %d iterations, final norm %.3e\n", name.c_str(), int(solver.GetConverged()),
                solver.GetNumIterations(), solver.GetFinalNorm());
  };
  {
    CGSolver cg;
    cg.SetRelTol(meta.at("rel_tol"));
    cg.SetMaxIter(int(meta.at("max_iter")));
    cg.SetOperator(*J);
    if (jacobi) cg.SetPreconditioner(prec);
    cg.iterative_mode = false;
    cg.Mult(B, x);
    report("cg", cg);
  }
  for (int kdim : {20, 50}) {
    GMRESSolver gmres;
    gmres.SetKDim(kdim);
    gmres.SetRelTol(meta.at("rel_tol"));
    gmres.SetMaxIter(int(meta.at("max_iter")));
    gmres.SetOperator(*J);
    if (jacobi) gmres.SetPreconditioner(prec);
    gmres.iterative_mode = false;
    gmres.Mult(B, x);
    report("gmres" + std::to_string(kdim), gmres);
    if (kdim == 20) {
      x                    = X;
      gmres.iterative_mode = true;
      gmres.Mult(B, x);
      report("gmres20_guess", gmres);
      gmres.iterative_mode = false;
      gmres.SetOperator(op);  // matrix-free over Gradient::Mult
      gmres.Mult(B, x);
      report("gmres20_mf", gmres);
    }
  }
  {
    GMRESSolver gmres;  // max_iter inside the second cycle
    gmres.SetKDim(5);
    gmres.SetRelTol(meta.at("rel_tol"));
    gmres.SetMaxIter(7);
    gmres.SetOperator(*J);
    if (jacobi) gmres.SetPreconditioner(prec);
    gmres.iterative_mode = false;
    gmres.Mult(B, x);
    report("gmres5_maxiter7", gmres);
  }
  return 0;
}

int run_j2(const std::map<std::string, double>& meta)
{
  Problem P = load_problem(meta, false);
  auto    q = read_bin<double>("params");  // E nu sigma_y Hi Hk density geom_nonlin
  using Material = solid_mechanics::J2SmallStrain<solid_mechanics::LinearHardening>;
  Material m{q[0], q[1], solid_mechanics::LinearHardening{q[2], q[3]}, q[4], q[5]};
  auto     gn = q[6] != 0.0 ? GeometricNonlinearities::On : GeometricNonlinearities::Off;
  SolidFunctional residual(&P.space, {&P.shape_space, &P.space, &P.space});
  const int nq    = num_quadrature_points(P.space.geometry, p + 1);
  auto      qdata = std::make_shared<QuadratureData<Material::State>>(P.mesh.GetNE(), nq, Material::State{});
  Domain    whole = EntireDomain(P.mesh);
  residual.AddDomainIntegral(Dimension<dim>{}, DependsOn<0, 1, 2>{}, SolidIntegrand<Material>(MaterialStressFunctor<Material>(m, gn)),
                             whole, qdata);
  Vector S(read_bin<double>("S")), U(read_bin<double>("U")), A(read_bin<double>("A")), dU(read_bin<double>("dU"));
  auto [r, drdu] = residual(0.0, S, differentiate_wrt(U), A);
  write_bin("r", r.to_host());
  write_bin("dr", drdu(dU).to_host());
  // the post-convergence pass of SolidMechanics (solid_mechanics.hpp:1282-1287): commit the state
  residual.updateQdata(true);
  residual(0.0, S, U, A);
  residual.updateQdata(false);
  auto                states = qdata->to_host();
  std::vector<double> flat(states.size() * sizeof(Material::State) / sizeof(double));
  std::memcpy(flat.data(), states.data(), flat.size() * sizeof(double));
  write_bin("qdata1", flat);
  write_bin("r_after_commit", residual(0.0, S, U, A).to_host());
  return 0;
}

/// finite-deformation J2 (solid_material.hpp:261-330) with a DEFAULT-constructed state buffer -- State{} must be the
/// virgin material (Fpinv = I) -- and the checkpoint round trip of the quadrature-point state (to_host / from_host: what a
/// restart does with the buffer the reference saves through sidre, state_manager.cpp:213-224)
int run_j2_finite(const std::map<std::string, double>& meta)
{
  Problem P = load_problem(meta, false);
  auto    q = read_bin<double>("params");  // E nu sigma_y Hi density geom_nonlin
  using Material = solid_mechanics::J2<solid_mechanics::LinearHardening>;
  Material m{q[0], q[1], solid_mechanics::LinearHardening{q[2], q[3]}, q[4]};
  auto     gn = q[5] != 0.0 ? GeometricNonlinearities::On : GeometricNonlinearities::Off;
  const int nq = num_quadrature_points(P.space.geometry, p + 1);
  Vector S(read_bin<double>("S")), U(read_bin<double>("U")), A(read_bin<double>("A")), dU(read_bin<double>("dU"));
  std::vector<Material::State> checkpoint;
  {
    SolidFunctional residual(&P.space, {&P.shape_space, &P.space, &P.space});
    auto            qdata = std::make_shared<QuadratureData<Material::State>>(P.mesh.GetNE(), nq);  // State{} everywhere
    Domain          whole = EntireDomain(P.mesh);
    residual.AddDomainIntegral(Dimension<dim>{}, DependsOn<0, 1, 2>{},
                               SolidIntegrand<Material>(MaterialStressFunctor<Material>(m, gn)), whole, qdata);
    auto [r, drdu] = residual(0.0, S, differentiate_wrt(U), A);
    write_bin("r", r.to_host());
    write_bin("dr", drdu(dU).to_host());
    residual.updateQdata(true);
    residual(0.0, S, U, A);
    residual.updateQdata(false);
    checkpoint = qdata->to_host();
    std::vector<double> flat(checkpoint.size() * sizeof(Material::State) / sizeof(double));
    std::memcpy(flat.data(), checkpoint.data(), flat.size() * sizeof(double));
    write_bin("qdata1", flat);
    write_bin("r_after_commit", residual(0.0, S, U, A).to_host());
  }
  {  // restart: a fresh Functional and state buffer, restored from the checkpoint
    SolidFunctional residual(&P.space, {&P.shape_space, &P.space, &P.space});
    auto            qdata = std::make_shared<QuadratureData<Material::State>>(P.mesh.GetNE(), nq);
    qdata->from_host(checkpoint);
    Domain whole = EntireDomain(P.mesh);
    residual.AddDomainIntegral(Dimension<dim>{}, DependsOn<0, 1, 2>{},
                               SolidIntegrand<Material>(MaterialStressFunctor<Material>(m, gn)), whole, qdata);
    write_bin("r_after_restart", residual(0.0, S, U, A).to_host());
  }
  return 0;
}

int run_pressure(const std::map<std::string, double>& meta)
{
  Problem P = load_problem(meta, true);
  auto    q = read_bin<double>("params");  // pressure geom_nonlin
  SolidFunctional residual(&P.space, {&P.shape_space, &P.space, &P.space});
  auto gn = q[1] != 0.0 ? GeometricNonlinearities::On : GeometricNonlinearities::Off;
  // SolidMechanics::setPressure (solid_mechanics.hpp:1075-1101) inside ShapeAwareFunctional::AddBoundaryIntegral
  using Integrand = ShapeAwareBoundaryIntegrandWrapper<PressureIntegrand<UniformPressure>>;
  residual.AddBoundaryIntegral(Dimension<dim - 1>{}, DependsOn<0, 1, 2>{},
                               Integrand(PressureIntegrand<UniformPressure>(UniformPressure{q[0]}, gn)), P.mesh);
  Vector S(read_bin<double>("S")), U(read_bin<double>("U")), A(read_bin<double>("A")), dU(read_bin<double>("dU")),
      dS(read_bin<double>("dS"));
  auto [r, drdu] = residual(0.0, S, differentiate_wrt(U), A);
  write_bin("r", r.to_host());
  write_bin("dr", drdu(dU).to_host());
  auto [r2, drds] = residual(0.0, differentiate_wrt(S), U, A);
  write_bin("dr_dshape", drds(dS).to_host());
  return 0;
}

/// a q-function that exists nowhere in the library: nonlinear reaction-diffusion with a position-dependent source
struct UserQFunction {
  double a, b;
  template <typename X, typename T>
  SERAC_HOST_DEVICE auto operator()(double t, X x, T temperature) const
  {
    auto u      = get<0>(temperature);
    auto du_dx  = get<1>(temperature);
    auto X0     = get<0>(x);
    auto source = a * u * u * u + (X0[0] + 2.0 * X0[1] - X0[2]) * (1.0 + t);
    auto flux   = (b + u * u) * du_dx;
    return serac::tuple{source, flux};
  }
};

int run_user(const std::map<std::string, double>& meta)
{
  Problem P = load_problem(meta, false);
  Functional<H1<p>(H1<p>)> residual(&P.scalar_space, {&P.scalar_space});
  Domain whole = EntireDomain(P.mesh);
  residual.AddDomainIntegral(Dimension<dim>{}, DependsOn<0>{}, UserQFunction{0.7, 1.3}, whole);
  const int64_t       n = P.scalar_space.GetVSize();
  std::vector<double> u(static_cast<size_t>(n)), du(static_cast<size_t>(n));
  for (int64_t i = 0; i < n; i++) {
    u[size_t(i)]  = 0.5 * std::sin(0.37 * double(i) + 0.1);
    du[size_t(i)] = std::cos(0.11 * double(i)) * (0.3 + double(i % 7) / 7.0);
  }
  Vector U(u), dU(du);
  auto [r, drdu]          = residual(0.5, differentiate_wrt(U));
  std::vector<double> r0  = r.to_host();
  std::vector<double> jvp = drdu(dU).to_host();
  auto                K   = assemble(drdu);
  // (1) J du from the assembled matrix
  double err_mat = 0.0, norm = 0.0;
  for (int64_t i = 0; i < n; i++) {
    double s = 0.0;
    for (int k = K->row_ptr[size_t(i)]; k < K->row_ptr[size_t(i) + 1]; k++) s += K->values[size_t(k)] * du[size_t(K->col_ind[size_t(k)])];
    err_mat += (s - jvp[size_t(i)]) * (s - jvp[size_t(i)]);
    norm += jvp[size_t(i)] * jvp[size_t(i)];
  }
  err_mat = std::sqrt(err_mat / norm);
  // (2) central finite difference of the residual
  const double        eps = 1e-5;
  std::vector<double> up(u), um(u);
  for (int64_t i = 0; i < n; i++) {
    up[size_t(i)] += eps * du[size_t(i)];
    um[size_t(i)] -= eps * du[size_t(i)];
  }
  std::vector<double> rp = residual(0.5, Vector(up)).to_host();
  std::vector<double> rm = residual(0.5, Vector(um)).to_host();
  double              err_fd = 0.0;
  for (int64_t i = 0; i < n; i++) {
    const double fd = (rp[size_t(i)] - rm[size_t(i)]) / (2.0 * eps);
    err_fd += (fd - jvp[size_t(i)]) * (fd - jvp[size_t(i)]);
  }
  err_fd = std::sqrt(err_fd / norm);
  // (3) the plain evaluation returns the same residual as the differentiated one
  std::vector<double> rplain = residual(0.5, U).to_host();
  double              err_r  = 0.0, rn = 0.0;
  for (int64_t i = 0; i < n; i++) {
    err_r += (rplain[size_t(i)] - r0[size_t(i)]) * (rplain[size_t(i)] - r0[size_t(i)]);
    rn += r0[size_t(i)] * r0[size_t(i)];
  }
  err_r = std::sqrt(err_r / rn);
  std::printf("user q-function: |J du|=%.6e  matrix-vs-action %.2e  finite-difference %.2e  plain-vs-AD residual %.2e\n",
              std::sqrt(norm), err_mat, err_fd, err_r);
  write_bin("r", r0);
  write_bin("dr", jvp);
  const bool ok = err_mat < 1e-12 && err_fd < 5e-8 && err_r < 1e-14 && norm > 0.0;
  std::printf(ok ? "PASS\n" : "FAIL\n");
  return ok ? 0 : 1;
}

/// the q-function of the reference's own solver test (src/serac/numerics/tests/equationsolver.cpp:66-73)
struct SineReactionDiffusion {
  template <typename X, typename T>
  SERAC_HOST_DEVICE auto operator()(double /*t*/, X /*x*/, T scalar) const
  {
    auto u      = get<0>(scalar);
    auto du_dx  = get<1>(scalar);
    auto source = 0.5 * sin(u);
    auto flux   = du_dx;
    return serac::tuple{source, flux};
  }
};

/// EquationSolverSuite of the reference (src/serac/numerics/tests/equationsolver.cpp:33-121), on this fixture's mesh:
/// solve r(x) - r(x_exact) = 0 from x = 0 with Newton (mfem::NewtonSolver's loop: stop at ||r|| <= max(rel ||r0||, abs))
/// and CG / GMRES + Jacobi (tolerances of the reference test), every step on the device; the reference's assertion is
/// |x_i - x_exact_i| / x_exact_i < 1e-6 for every dof.
int run_equation_solver(const std::map<std::string, double>& meta)
{
  Problem P = load_problem(meta, false);
  Functional<H1<p>(H1<p>)> residual(&P.scalar_space, {&P.scalar_space});
  Domain whole = EntireDomain(P.mesh);
  residual.AddDomainIntegral(Dimension<dim>{}, DependsOn<0>{}, SineReactionDiffusion{}, whole);
  const int64_t       n = P.scalar_space.GetVSize();
  std::vector<double> xe(static_cast<size_t>(n));
  unsigned long long  lcg = 12345;  // x_exact.Randomize(0): uniform values; kept away from 0 for the relative test
  for (auto& v : xe) {
    lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL;
    v   = 0.1 + 0.9 * double(lcg >> 11) / double(1ULL << 53);
  }
  Vector x_exact(xe);
  Vector r_exact = residual(0.0, x_exact);  // copy: the Functional's output vector is reused by the next call
  DofList no_ess{std::vector<int>{}};
  bool    all_ok = true;
  for (int which = 0; which < 4; which++) {  // CG, GMRES (assembled), GMRES (matrix-free), GMRES + l1-Jacobi (HeatTransfer's default)
    Vector x(n), dx(n), r(n);
    x = 0.0;
    CGSolver    cg;
    GMRESSolver gmres;
    cg.SetRelTol(1.0e-10);
    cg.SetAbsTol(1.0e-12);
    cg.SetMaxIter(500);
    cg.iterative_mode = false;
    gmres.SetRelTol(1.0e-10);
    gmres.SetAbsTol(1.0e-12);
    gmres.SetMaxIter(500);
    gmres.iterative_mode = false;
    double norm0 = 0.0, norm = 0.0;
    int    it = 0, linear_its = 0;
    for (;; it++) {
      auto [val, grad] = residual(0.0, differentiate_wrt(x));
      r                = val;
      Add(r, -1.0, r_exact);  // r -= residual(x_exact)
      norm = Norml2(r);
      if (it == 0) norm0 = norm;
      if (norm <= std::max(1.0e-10 * norm0, 1.0e-12) || it >= 100) break;
      auto             J = assemble_on_device(grad);
      JacobiSmoother   prec(*J);
      L1JacobiSmoother prec_l1(*J);
      if (which == 0) {
        cg.SetOperator(*J);
        cg.SetPreconditioner(prec);
        cg.Mult(r, dx);
        linear_its += cg.GetNumIterations();
        all_ok = all_ok && cg.GetConverged();
      } else {
        GradientOperator op(grad, no_ess);
        if (which == 2)
          gmres.SetOperator(op);
        else
          gmres.SetOperator(*J);
        if (which == 3)
          gmres.SetPreconditioner(prec_l1);
        else
          gmres.SetPreconditioner(prec);
        gmres.Mult(r, dx);
        linear_its += gmres.GetNumIterations();
        all_ok = all_ok && gmres.GetConverged();
      }
      Add(x, -1.0, dx);
    }
    std::vector<double> xc = x.to_host();
    double              worst = 0.0;
    for (size_t i = 0; i < xc.size(); i++) worst = std::max(worst, std::abs(xc[i] - xe[i]) / xe[i]);
    const bool ok = worst < 1.0e-6 && it < 100;
    std::printf("equation solver [%s]: %d Newton steps, %d linear iterations, |r| %.3e -> %.3e, worst relative error %.3e %s\n",
                which == 0 ? "CG + Jacobi" : (which == 1 ? "GMRES + Jacobi" : (which == 2 ? "GMRES + Jacobi, matrix-free" : "GMRES + l1-Jacobi")), it, linear_its,
                norm0, norm, worst, ok ? "PASS" : "FAIL");
    all_ok = all_ok && ok;
    write_bin(std::string("x_") + (which == 0 ? "cg" : (which == 1 ? "gmres" : (which == 2 ? "gmres_mf" : "gmres_l1"))), xc);
  }
  write_bin("x_exact", xe);
  return all_ok ? 0 : 1;
}

/// quantity of interest defined in this file: the integrand of the library's `qoi_quadratic_*` (and of the oracle kind)
struct QuadraticQoi {
  double a, b;
  template <typename X, typename U>
  SERAC_HOST_DEVICE auto operator()(double, X x, U u) const
  {
    auto uv = get<0>(u);
    auto du = get<1>(u);
    auto X0 = get<0>(x);
    return a * (1.0 + X0[0]) * dot(uv, uv) + (0.5 * b) * inner(du, du) + ::sin(X0[1]);
  }
};

int run_qoi(const std::map<std::string, double>& meta)
{
  Problem P = load_problem(meta, false);
  auto    q = read_bin<double>("params");
  Functional<double(H1<p, dim>)> qoi({&P.space});
  Domain whole = EntireDomain(P.mesh);
  qoi.AddDomainIntegral(Dimension<dim>{}, DependsOn<0>{}, QuadraticQoi{q[0], q[1]}, whole);
  Vector U(read_bin<double>("U")), dU(read_bin<double>("dU"));
  const double plain = qoi(0.0, U);
  auto [value, dqdu] = qoi(0.0, differentiate_wrt(U));
  write_bin("value", std::vector<double>{plain, value});
  write_bin("dvalue", dqdu(dU).to_host());
  write_bin("gradient", gradient_vector(dqdu));
  return 0;
}

struct QuadraticBoundaryQoi {
  double a, b;
  template <typename X, typename U>
  SERAC_HOST_DEVICE auto operator()(double, X x, U u) const
  {
    auto uv = get<0>(u);
    auto du = get<1>(u);
    auto X0 = get<0>(x);
    return a * (1.0 + X0[0]) * dot(uv, uv) + (0.5 * b) * inner(du, du);
  }
};

int run_boundary_qoi(const std::map<std::string, double>& meta)
{
  Problem P = load_problem(meta, true);
  auto    q = read_bin<double>("params");
  Functional<double(H1<p, dim>)> qoi({&P.space});
  qoi.AddBoundaryIntegral(Dimension<dim - 1>{}, DependsOn<0>{}, QuadraticBoundaryQoi{q[0], q[1]}, P.mesh);
  Vector U(read_bin<double>("U")), dU(read_bin<double>("dU"));
  auto [value, dqdu] = qoi(0.0, differentiate_wrt(U));
  write_bin("value", std::vector<double>{value});
  write_bin("dvalue", dqdu(dU).to_host());
  write_bin("gradient", gradient_vector(dqdu));
  return 0;
}

int main(int argc, char** argv)
{
  if (argc != 3) {
    std::fprintf(stderr, "usage: functional_test <solid|j2|pressure|user> <directory>\n");
    return 2;
  }
  const std::string which = argv[1];
  g_dir                   = argv[2];
  auto meta               = read_meta();
  if (int(meta.at("p")) != p || int(meta.at("dim")) != dim) {
    std::fprintf(stderr, "this driver is compiled for p = %d, dim = %d\n", p, dim);
    return 2;
  }
  int rc = 2;
  if (which == "solid") rc = run_solid(meta);
  if (which == "j2") rc = run_j2(meta);
  if (which == "j2_finite") rc = run_j2_finite(meta);
  if (which == "pressure") rc = run_pressure(meta);
  if (which == "user") rc = run_user(meta);
  if (which == "newton") rc = run_newton(meta, false) || run_newton(meta, true) || run_newton(meta, false, true);
  if (which == "krylov") rc = run_krylov(meta);
  if (which == "equation_solver") rc = run_equation_solver(meta);
  if (which == "qoi") rc = run_qoi(meta);
  if (which == "boundary_qoi") rc = run_boundary_qoi(meta);
  std::printf("launches: %lld\n", sb_launch_count());
  return rc;
}
