// Self-contained check of the solver layer of the C ABI (include/serac_b200_solvers.h) on a problem small enough to
// restate in this file: the stiffness matrix of a chain of two-node elements with a varying coefficient, assembled
// through sb_csr_create / sb_csr_assemble, then essential-dof elimination, EliminateBC, SpMV, Jacobi and CG -- each
// compared with a plain host loop written here (test code; the product has no CPU path).  No Python, no fixtures:
//   krylov_test [num_elements]        exit code 0 = all checks passed
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "serac_b200_solvers.h"

#define CK(expr)                                                                          \
  do {                                                                                    \
    int _rc = (expr);                                                                     \
    if (_rc != 0) {                                                                       \
      std::fprintf(stderr, "%s failed (%d): %s\n", #expr, _rc, sb_last_error());          \
      return 2;                                                                           \
    }                                                                                     \
  } while (0)
#define CU(expr)                                                                          \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      std::fprintf(stderr, "%s: %s\n", #expr, cudaGetErrorString(_e));                    \
      return 2;                                                                           \
    }                                                                                     \
  } while (0)

static int g_failures = 0;
static void expect(bool ok, const char* what, double got = 0, double bound = 0)
{
  std::printf("%s %s (%.3e vs %.3e)\n", ok ? "PASS" : "FAIL", what, got, bound);
  if (!ok) g_failures++;
}

template <class T>
T* to_device(const std::vector<T>& v)
{
  T* p = nullptr;
  cudaMalloc(reinterpret_cast<void**>(&p), (v.size() + 1) * sizeof(T));
  cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  return p;
}
template <class T>
std::vector<T> to_host(const T* p, size_t n)
{
  std::vector<T> v(n);
  cudaMemcpy(v.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost);
  return v;
}

static double rel_l2(const std::vector<double>& a, const std::vector<double>& b)
{
  double num = 0, den = 0;
  for (size_t i = 0; i < a.size(); i++) {
    num += (a[i] - b[i]) * (a[i] - b[i]);
    den += b[i] * b[i];
  }
  return std::sqrt(num / (den > 0 ? den : 1));
}

struct HostCsr {
  std::vector<int>    rp, ci;
  std::vector<double> v;
  std::vector<double> mult(const std::vector<double>& x) const
  {
    std::vector<double> y(rp.size() - 1, 0.0);
    for (size_t i = 0; i + 1 < rp.size(); i++)
      for (int j = rp[i]; j < rp[i + 1]; j++) y[i] += v[size_t(j)] * x[size_t(ci[size_t(j)])];
    return y;
  }
};

/// mfem::CGSolver::Mult with a Jacobi preconditioner, restated for the check (mfem/linalg/solvers.cpp:719-897)
static int host_cg(const HostCsr& A, const std::vector<double>& dinv, const std::vector<double>& b, std::vector<double>& x,
                   double rel_tol, int max_iter, bool* converged)
{
  const size_t        n = b.size();
  std::vector<double> r = b, z(n), d(n);
  x.assign(n, 0.0);
  auto dot = [&](const std::vector<double>& p, const std::vector<double>& q) {
    double s = 0;
    for (size_t i = 0; i < n; i++) s += p[i] * q[i];
    return s;
  };
  for (size_t i = 0; i < n; i++) d[i] = z[i] = dinv[i] * r[i];
  double nom = dot(d, r);
  const double r0 = nom * rel_tol * rel_tol;
  *converged      = false;
  if (nom <= r0) {
    *converged = true;
    return 0;
  }
  z          = A.mult(d);
  double den = dot(z, d);
  if (den == 0.0) return 0;
  for (int i = 1;;) {
    const double alpha = nom / den;
    for (size_t k = 0; k < n; k++) {
      x[k] += alpha * d[k];
      r[k] -= alpha * z[k];
      z[k] = dinv[k] * r[k];
    }
    const double betanom = dot(r, z);
    if (betanom <= r0) {
      *converged = true;
      return i;
    }
    if (++i > max_iter) return max_iter;
    const double beta = betanom / nom;
    for (size_t k = 0; k < n; k++) d[k] = z[k] + beta * d[k];
    z   = A.mult(d);
    den = dot(d, z);
    if (den == 0.0) return i;
    nom = betanom;
  }
}

struct SpmvCtx {
  sb_csr_t      A;
  const double* values;
};
static int spmv_callback(void* ctx, const double* x, double* y, void* stream)
{
  auto* c = static_cast<SpmvCtx*>(ctx);
  return sb_csr_spmv(c->A, c->values, x, y, stream);
}

int main(int argc, char** argv)
{
  const int ne = argc > 1 ? std::atoi(argv[1]) : 3000;
  const int n  = ne + 1;
  const size_t N = static_cast<size_t>(n);
  // chain of two-node elements; DoF words: index in bits 0-47 (F/element_restriction.hpp:37-43)
  std::vector<uint64_t> dof_info(size_t(ne) * 2);
  for (int e = 0; e < ne; e++) {
    dof_info[size_t(2 * e)]     = uint64_t(e);
    dof_info[size_t(2 * e + 1)] = uint64_t(e + 1);
  }
  sb_restriction_t R = nullptr;
  CK(sb_restriction_create(dof_info.data(), ne, 2, 1, n, SB_ORDERING_BYVDIM, &R));
  sb_csr_t A = nullptr;
  CK(sb_csr_create(R, R, &A));
  const int64_t nnz = sb_csr_nnz(A);
  expect(sb_csr_num_rows(A) == n && nnz == 3 * int64_t(n) - 2, "tridiagonal pattern", double(nnz), 3.0 * n - 2);
  // element matrices k [[1,-1],[-1,1]] + m I (reaction term)
  std::vector<double> K_e(size_t(ne) * 4);
  for (int e = 0; e < ne; e++) {
    const double k         = 1.0 + 0.5 * std::sin(0.37 * e);
    const double m         = 0.05;  // condition number ~ 1e2: CG converges in ~1e2 steps whatever the sum order
    K_e[size_t(4 * e)]     = k + m;
    K_e[size_t(4 * e + 1)] = -k;
    K_e[size_t(4 * e + 2)] = -k;
    K_e[size_t(4 * e + 3)] = k + m;
  }
  double* K_e_dev  = to_device(K_e);
  double* vals_dev = nullptr;
  CU(cudaMalloc(reinterpret_cast<void**>(&vals_dev), size_t(nnz) * sizeof(double)));
  CK(sb_csr_assemble(A, K_e_dev, vals_dev, nullptr));
  HostCsr H;
  H.rp.resize(N + 1);
  H.ci.resize(size_t(nnz));
  CK(sb_csr_pattern_to_host(A, H.rp.data(), H.ci.data()));
  CU(cudaDeviceSynchronize());
  H.v = to_host(vals_dev, size_t(nnz));

  // essential dofs: both ends and a few inside
  std::vector<int> ess = {0, n / 3, n / 2, n - 1};
  int*             ess_dev = to_device(ess);
  std::vector<double> X(N), B(N), V(N);
  for (int i = 0; i < n; i++) {
    X[size_t(i)] = std::cos(0.11 * i);
    B[size_t(i)] = std::sin(0.05 * i) + 0.3;
    V[size_t(i)] = 0.5 + 0.5 * std::sin(0.7 * i);
  }
  double *X_dev = to_device(X), *B_dev = to_device(B), *V_dev = to_device(V);

  // --- elimination: bit-exact against the host loop -----------------------------------------------------------------
  std::vector<char> mark(N, 0);
  for (int b : ess) mark[size_t(b)] = 1;
  HostCsr He = H, Hae = H;
  for (int i = 0; i < n; i++)
    for (int j = H.rp[size_t(i)]; j < H.rp[size_t(i) + 1]; j++) {
      const int  c    = H.ci[size_t(j)];
      const bool elim = mark[size_t(i)] || mark[size_t(c)];
      const double keep = (c == i) ? 1.0 : 0.0;
      He.v[size_t(j)]   = elim ? keep : H.v[size_t(j)];
      Hae.v[size_t(j)]  = elim ? H.v[size_t(j)] - keep : 0.0;
    }
  double* vale_dev = nullptr;
  CU(cudaMalloc(reinterpret_cast<void**>(&vale_dev), size_t(nnz) * sizeof(double)));
  CK(sb_csr_eliminate_rows_cols(A, vals_dev, ess_dev, int64_t(ess.size()), vale_dev, nullptr));
  CU(cudaDeviceSynchronize());
  expect(to_host(vals_dev, size_t(nnz)) == He.v, "eliminated matrix bit-exact");
  expect(to_host(vale_dev, size_t(nnz)) == Hae.v, "eliminated part Ae bit-exact");

  // --- diagonal, SpMV, EliminateBC -------------------------------------------------------------------------------------
  double *diag_dev = nullptr, *dinv_dev = nullptr, *y_dev = nullptr, *x_dev = nullptr;
  CU(cudaMalloc(reinterpret_cast<void**>(&diag_dev), size_t(n) * sizeof(double)));
  CU(cudaMalloc(reinterpret_cast<void**>(&dinv_dev), size_t(n) * sizeof(double)));
  CU(cudaMalloc(reinterpret_cast<void**>(&y_dev), size_t(n) * sizeof(double)));
  CU(cudaMalloc(reinterpret_cast<void**>(&x_dev), size_t(n) * sizeof(double)));
  CK(sb_csr_diagonal(A, vals_dev, diag_dev, nullptr));
  CK(sb_jacobi_setup(diag_dev, dinv_dev, n, nullptr));
  CK(sb_csr_spmv(A, vals_dev, V_dev, y_dev, nullptr));
  CU(cudaDeviceSynchronize());
  std::vector<double> diag_ref(N), dinv_ref(N);
  for (int i = 0; i < n; i++)
    for (int j = H.rp[size_t(i)]; j < H.rp[size_t(i) + 1]; j++)
      if (H.ci[size_t(j)] == i) diag_ref[size_t(i)] = He.v[size_t(j)];
  for (int i = 0; i < n; i++) dinv_ref[size_t(i)] = 1.0 / diag_ref[size_t(i)];
  expect(to_host(diag_dev, size_t(n)) == diag_ref, "diagonal bit-exact");
  expect(rel_l2(to_host(dinv_dev, size_t(n)), dinv_ref) < 1e-15, "Jacobi D^-1");
  double e = rel_l2(to_host(y_dev, size_t(n)), He.mult(V));
  expect(e < 1e-14, "SpMV", e, 1e-14);
  CK(sb_csr_eliminate_bc(A, vals_dev, vale_dev, ess_dev, int64_t(ess.size()), X_dev, B_dev, nullptr));
  CU(cudaDeviceSynchronize());
  std::vector<double> Bbc = B, AeX = Hae.mult(X);
  for (int i = 0; i < n; i++) Bbc[size_t(i)] -= AeX[size_t(i)];
  for (int b : ess) Bbc[size_t(b)] = X[size_t(b)];
  e = rel_l2(to_host(B_dev, size_t(n)), Bbc);
  expect(e < 1e-14, "EliminateBC", e, 1e-14);
  CK(sb_vector_set_subvector(y_dev, ess_dev, int64_t(ess.size()), 0.0, nullptr));
  CU(cudaDeviceSynchronize());
  {
    auto y  = to_host(y_dev, size_t(n));
    bool ok = true;
    for (int b : ess) ok = ok && y[size_t(b)] == 0.0;
    expect(ok, "SetSubVector");
  }

  // --- CG --------------------------------------------------------------------------------------------------------------
  const double rel_tol  = 1e-10;
  const int    max_iter = 20 * n;
  std::vector<double> x_ref;
  bool                conv_ref = false;
  const int           it_ref   = host_cg(He, dinv_ref, Bbc, x_ref, rel_tol, max_iter, &conv_ref);
  sb_cg_t S = nullptr;
  CK(sb_cg_create(n, &S));
  sb_cg_options opt{rel_tol, 0.0, max_iter, 0, 1};
  sb_cg_result  res{};
  CK(sb_cg_solve_csr(S, A, vals_dev, dinv_dev, B_dev, x_dev, &opt, &res, nullptr));
  CU(cudaDeviceSynchronize());
  auto x1 = to_host(x_dev, size_t(n));
  std::printf("CG: host %d iterations (converged %d), device %d iterations (converged %d), final norm %.3e / initial %.3e\n",
              it_ref, int(conv_ref), res.final_iter, res.converged, res.final_norm, res.initial_norm);
  expect(res.converged == 1 && conv_ref, "CG converged");
  expect(std::abs(res.final_iter - it_ref) <= std::max(3, it_ref / 50), "CG iteration count", res.final_iter, it_ref);
  e = rel_l2(x1, x_ref);
  expect(e < 1e-6, "CG solution vs host CG", e, 1e-6);
  {
    auto   Ax  = He.mult(x1);
    double num = 0, den = 0;
    for (int i = 0; i < n; i++) {
      num += (Ax[size_t(i)] - Bbc[size_t(i)]) * (Ax[size_t(i)] - Bbc[size_t(i)]);
      den += Bbc[size_t(i)] * Bbc[size_t(i)];
    }
    expect(std::sqrt(num / den) < 1e-7, "CG residual ||Ax - b|| / ||b||", std::sqrt(num / den), 1e-7);
  }
  // the host running ahead of the verdict changes nothing
  opt.check_every = 16;
  sb_cg_result res16{};
  CK(sb_cg_solve_csr(S, A, vals_dev, dinv_dev, B_dev, x_dev, &opt, &res16, nullptr));
  CU(cudaDeviceSynchronize());
  expect(res16.final_iter == res.final_iter && res16.converged == res.converged && to_host(x_dev, size_t(n)) == x1,
         "check_every = 16: same iterates bit for bit", res16.final_iter, res.final_iter);
  // operator form (callback) == CSR form up to the order of the (d, z) sum
  SpmvCtx       ctx{A, vals_dev};
  sb_cg_result  resop{};
  opt.check_every = 1;
  CK(sb_cg_solve_operator(S, spmv_callback, &ctx, dinv_dev, B_dev, x_dev, &opt, &resop, nullptr));
  CU(cudaDeviceSynchronize());
  e = rel_l2(to_host(x_dev, size_t(n)), x1);
  expect(resop.converged == 1 && e < 1e-6, "operator form", e, 1e-6);
  // max_iter exit and iterative_mode
  sb_cg_options opt5{rel_tol, 0.0, 5, 0, 1};
  sb_cg_result  res5{};
  CK(sb_cg_solve_csr(S, A, vals_dev, dinv_dev, B_dev, x_dev, &opt5, &res5, nullptr));
  bool                conv5 = true;
  std::vector<double> x5_ref;
  host_cg(He, dinv_ref, Bbc, x5_ref, rel_tol, 5, &conv5);
  CU(cudaDeviceSynchronize());
  e = rel_l2(to_host(x_dev, size_t(n)), x5_ref);
  expect(res5.converged == 0 && res5.final_iter == 5 && !conv5 && e < 1e-12, "max_iter = 5 exit", e, 1e-12);
  sb_cg_options optg{rel_tol, 0.0, max_iter, 1, 1};
  sb_cg_result  resg{};
  CU(cudaMemcpy(x_dev, X.data(), size_t(n) * sizeof(double), cudaMemcpyHostToDevice));
  CK(sb_cg_solve_csr(S, A, vals_dev, dinv_dev, B_dev, x_dev, &optg, &resg, nullptr));
  CU(cudaDeviceSynchronize());
  e = rel_l2(to_host(x_dev, size_t(n)), x1);
  expect(resg.converged == 1 && e < 1e-6, "iterative_mode from a guess reaches the same solution", e, 1e-6);

  // --- GMRES(m): same system, so the same solution; restarts (m = 10) and mfem's default m = 50 -------------------------
  for (int kdim : {10, 50}) {
    sb_gmres_t G = nullptr;
    CK(sb_gmres_create(n, kdim, &G));
    sb_cg_options og{rel_tol, 0.0, max_iter, 0, 1};
    sb_cg_result  rg{}, rg2{};
    CK(sb_gmres_solve_csr(G, A, vals_dev, dinv_dev, B_dev, x_dev, &og, &rg, nullptr));
    CU(cudaDeviceSynchronize());
    auto xg = to_host(x_dev, N);
    e       = rel_l2(xg, x1);
    std::printf("GMRES(%d): %d iterations (converged %d), final %.3e / initial %.3e\n", kdim, rg.final_iter, rg.converged,
                rg.final_norm, rg.initial_norm);
    expect(rg.converged == 1 && e < 1e-6, "GMRES solution vs CG solution", e, 1e-6);
    expect(rg.final_norm <= rel_tol * rg.initial_norm * (1 + 1e-12), "GMRES stopping rule", rg.final_norm,
           rel_tol * rg.initial_norm);
    {  // the reported residual estimate is honest: ||D^-1 (b - A x)|| within a factor of the estimate
      auto   Ax = He.mult(xg);
      double tr = 0;
      for (size_t i = 0; i < N; i++) tr += std::pow(dinv_ref[i] * (Bbc[i] - Ax[i]), 2);
      expect(std::sqrt(tr) < 10 * rel_tol * rg.initial_norm, "GMRES true preconditioned residual", std::sqrt(tr),
             10 * rel_tol * rg.initial_norm);
    }
    CK(sb_gmres_solve_csr(G, A, vals_dev, dinv_dev, B_dev, x_dev, &og, &rg2, nullptr));
    CU(cudaDeviceSynchronize());
    expect(rg2.final_iter == rg.final_iter && to_host(x_dev, N) == xg, "GMRES is deterministic");
    SpmvCtx      ctxg{A, vals_dev};
    sb_cg_result rgo{};
    CK(sb_gmres_solve_operator(G, spmv_callback, &ctxg, dinv_dev, B_dev, x_dev, &og, &rgo, nullptr));
    CU(cudaDeviceSynchronize());
    expect(rgo.final_iter == rg.final_iter && to_host(x_dev, N) == xg, "GMRES operator form = CSR form bit for bit");
    const int     cap = (kdim == 10) ? 13 : 20;  // inside the second cycle / inside the first
    sb_cg_options om{rel_tol, 0.0, cap, 0, 1};
    sb_cg_result  rm{};
    CK(sb_gmres_solve_csr(G, A, vals_dev, dinv_dev, B_dev, x_dev, &om, &rm, nullptr));
    CU(cudaDeviceSynchronize());
    expect(rm.converged == 0 && rm.final_iter == cap && rm.final_norm < rm.initial_norm, "GMRES max_iter exit",
           rm.final_iter, cap);
    CK(sb_gmres_destroy(G));
  }

  CK(sb_cg_destroy(S));
  CK(sb_csr_destroy(A));
  CK(sb_restriction_destroy(R));
  std::printf("%s: %d failure(s)\n", g_failures ? "FAILED" : "ALL PASSED", g_failures);
  return g_failures ? 1 : 0;
}
