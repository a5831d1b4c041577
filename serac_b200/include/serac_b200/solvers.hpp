// C++ host side of include/serac_b200_solvers.h: the objects a Newton step of SolidMechanics / HeatTransfer uses between
// two Functional evaluations, with the reference stack's names and call shapes, all device-resident:
//
//   reference (mfem / serac)                                    here
//   ---------------------------------------------------------   -------------------------------------------------------
//   mfem::Array<int> (essential true dofs)                      serac::DofList
//   mfem::HypreParMatrix from assemble(drdu)                    serac::DeviceCsrMatrix from assemble_on_device(drdu)
//   A.EliminateRowsCols(ess) -> Ae  (hypre_parcsr.cpp:476)      A.EliminateRowsCols(ess) -> Ae
//   mfem::EliminateBC(A, Ae, ess, X, B)  (hypre.cpp:3383)       serac::EliminateBC(A, Ae, ess, X, B)
//   r.SetSubVector(ess, 0.0)  (solid_mechanics.hpp:1129)        serac::SetSubVector(r, ess, 0.0)
//   mfem::OperatorJacobiSmoother / HypreSmoother Jacobi         serac::JacobiSmoother
//   mfem::CGSolver (solvers.cpp:719-897)                        serac::CGSolver (SetOperator / SetPreconditioner / Mult)
//   mfem::GMRESSolver (solvers.cpp:981-1165)                    serac::GMRESSolver (SetKDim, same calls)
//   Functional::Gradient as the operator (matrix-free)          serac::GradientOperator
//
// Failures abort with the library's message, as the reference does through SLIC_ERROR_ROOT.
#pragma once

#include <cmath>
#include <memory>
#include <vector>

#include "serac_b200/functional.hpp"
#include "serac_b200_solvers.h"

namespace serac {

/// mfem::Array<int> with device memory
class DofList {
 public:
  DofList() = default;
  explicit DofList(const std::vector<int>& host) : n_(int64_t(host.size()))
  {
    if (n_) {
      SERAC_B200_CUDA(cudaMalloc(&d_, sizeof(int) * size_t(n_)));
      SERAC_B200_CUDA(cudaMemcpy(d_, host.data(), sizeof(int) * size_t(n_), cudaMemcpyHostToDevice));
    }
  }
  DofList(const DofList&)            = delete;
  DofList& operator=(const DofList&) = delete;
  ~DofList() { cudaFree(d_); }
  int64_t    Size() const { return n_; }
  const int* Read() const { return d_; }

 private:
  int64_t n_ = 0;
  int*    d_ = nullptr;
};

inline void SetSubVector(Vector& v, const DofList& dofs, double value)
{
  SERAC_B200_CHECK(sb_vector_set_subvector(v.Write(), dofs.Read(), dofs.Size(), value, nullptr));
}
/// y += a x  (mfem::Vector::Add)
inline void Add(Vector& y, double a, const Vector& x)
{
  SERAC_B200_CHECK(sb_vector_axpy(y.Write(), a, x.Read(), y.Size(), nullptr));
}
inline double Dot(const Vector& x, const Vector& y)
{
  double s = 0.0;
  SERAC_B200_CHECK(sb_vector_dot(x.Read(), y.Read(), x.Size(), &s, nullptr));
  return s;
}
inline double Norml2(const Vector& x) { return std::sqrt(Dot(x, x)); }

/// the assembled Jacobian on the device: the pattern object belongs to the Functional (as the reference's lookup tables
/// own the CSR graph, functional.hpp:571-579), the values to this matrix
class DeviceCsrMatrix {
 public:
  DeviceCsrMatrix(sb_csr_t pattern, int64_t width) : pattern_(pattern), width_(width), values_(sb_csr_nnz(pattern)) {}
  int64_t  Height() const { return sb_csr_num_rows(pattern_); }
  int64_t  Width() const { return width_; }
  sb_csr_t pattern() const { return pattern_; }
  Vector&       values() { return values_; }
  const Vector& values() const { return values_; }
  /// mfem::Operator::Mult
  void Mult(const Vector& x, Vector& y) const
  {
    SERAC_B200_CHECK(sb_csr_spmv(pattern_, values_.Read(), x.Read(), y.Write(), nullptr));
  }
  /// HypreParMatrix::EliminateRowsCols: this -> [A_ii 0; 0 I], returns Ae = [0 A_ib; A_bi A_bb - I]
  std::unique_ptr<DeviceCsrMatrix> EliminateRowsCols(const DofList& ess)
  {
    auto Ae = std::make_unique<DeviceCsrMatrix>(pattern_, width_);
    SERAC_B200_CHECK(sb_csr_eliminate_rows_cols(pattern_, values_.Write(), ess.Read(), ess.Size(), Ae->values_.Write(),
                                                nullptr));
    return Ae;
  }
  void GetDiag(Vector& d) const
  {
    d.SetSize(Height());
    SERAC_B200_CHECK(sb_csr_diagonal(pattern_, values_.Read(), d.Write(), nullptr));
  }
  /// HypreParMatrix::Transpose() (the adjoint solve of SolidMechanics::reverseAdjointTimestep, solid_mechanics.hpp:1362-1371):
  /// a matrix of its own, on the transposed pattern (created on first use and kept with this matrix)
  std::unique_ptr<DeviceCsrMatrix> Transpose()
  {
    if (!transposed_pattern_) SERAC_B200_CHECK(sb_csr_transpose(pattern_, &transposed_pattern_));
    auto T = std::make_unique<DeviceCsrMatrix>(transposed_pattern_, Height());
    SERAC_B200_CHECK(sb_csr_transpose_values(transposed_pattern_, values_.Read(), T->values_.Write(), nullptr));
    return T;
  }
  /// mfem::Operator::MultTranspose (sensitivities: drdparam_mat->MultTranspose(adjoint, sensitivity), :1404-1424)
  void MultTranspose(const Vector& x, Vector& y)
  {
    auto T = Transpose();
    T->Mult(x, y);
  }
  ~DeviceCsrMatrix()
  {
    if (transposed_pattern_) sb_csr_destroy(transposed_pattern_);
  }
  DeviceCsrMatrix(const DeviceCsrMatrix&)            = delete;
  DeviceCsrMatrix& operator=(const DeviceCsrMatrix&) = delete;

 private:
  sb_csr_t pattern_;
  int64_t  width_;
  Vector   values_;
  sb_csr_t transposed_pattern_ = nullptr;  ///< owned; the transposed matrices made from it must not outlive this one
};

/// assemble(Gradient&) (functional.hpp:569-667) with the values left on the device
template <typename GradientT>
std::unique_ptr<DeviceCsrMatrix> assemble_on_device(GradientT& g)
{
  sb_csr_t A   = g.pattern();
  auto     out = std::make_unique<DeviceCsrMatrix>(A, g.Width());
  g.assemble_values(A, out->values().Write());
  return out;
}

inline void EliminateBC(const DeviceCsrMatrix& A, const DeviceCsrMatrix& Ae, const DofList& ess, const Vector& X, Vector& B)
{
  SERAC_B200_CHECK(sb_csr_eliminate_bc(A.pattern(), A.values().Read(), Ae.values().Read(), ess.Read(), ess.Size(),
                                       X.Read(), B.Write(), nullptr));
}

/// z = D^-1 r  (OperatorJacobiSmoother with damping 1, mfem/linalg/solvers.cpp:262-330)
class JacobiSmoother {
 public:
  explicit JacobiSmoother(const DeviceCsrMatrix& A)
  {
    Vector d;
    A.GetDiag(d);
    dinv_.SetSize(d.Size());
    SERAC_B200_CHECK(sb_jacobi_setup(d.Read(), dinv_.Write(), d.Size(), nullptr));
  }
  /// from a given scaling vector d: z = r / d
  explicit JacobiSmoother(const Vector& d)
  {
    dinv_.SetSize(d.Size());
    SERAC_B200_CHECK(sb_jacobi_setup(d.Read(), dinv_.Write(), d.Size(), nullptr));
  }
  const double* dinv() const { return dinv_.Read(); }

 private:
  Vector dinv_;
};

/// z = r / (sum_j |a_ij|): one sweep of hypre's l1-scaled Jacobi from a zero guess (mfem::HypreSmoother::l1Jacobi,
/// serac's Preconditioner::HypreL1Jacobi; include/serac_b200_solvers.h: sb_csr_row_l1_norms)
class L1JacobiSmoother : public JacobiSmoother {
 public:
  explicit L1JacobiSmoother(const DeviceCsrMatrix& A) : JacobiSmoother(l1_of(A)) {}

 private:
  static Vector l1_of(const DeviceCsrMatrix& A)
  {
    Vector l1(A.Height());
    SERAC_B200_CHECK(sb_csr_row_l1_norms(A.pattern(), A.values().Read(), l1.Write(), nullptr));
    return l1;
  }
};

/// matrix-free [J_ii 0; 0 I] over Functional::Gradient (Gradient::Mult, functional.hpp:556-559)
class GradientOperator {
 public:
  template <typename GradientT>
  GradientOperator(GradientT& g, const DofList& ess) : n_(g.Width())
  {
    SERAC_B200_CHECK(sb_functional_gradient_operator_create(g.functional_handle(), g.argument(), g.Width(), ess.Read(),
                                                            ess.Size(), &op_));
  }
  GradientOperator(const GradientOperator&)            = delete;
  GradientOperator& operator=(const GradientOperator&) = delete;
  ~GradientOperator() { sb_functional_gradient_operator_destroy(op_); }
  int64_t Height() const { return n_; }
  void    Mult(const Vector& x, Vector& y) const
  {
    SERAC_B200_CHECK(sb_functional_gradient_operator_apply(op_, x.Read(), y.Write(), nullptr));
  }
  sb_gradient_operator_t handle() const { return op_; }

 private:
  int64_t                n_;
  sb_gradient_operator_t op_ = nullptr;
};

/// mfem::CGSolver
class CGSolver {
 public:
  CGSolver() = default;
  CGSolver(const CGSolver&)            = delete;
  CGSolver& operator=(const CGSolver&) = delete;
  ~CGSolver() { sb_cg_destroy(S_); }
  void SetRelTol(double t) { opt_.rel_tol = t; }
  void SetAbsTol(double t) { opt_.abs_tol = t; }
  void SetMaxIter(int n) { opt_.max_iter = n; }
  /// how often the host looks at the device-side verdict (1 = every iteration, the reference's behaviour)
  void SetCheckEvery(int k) { opt_.check_every = k; }
  void SetOperator(const DeviceCsrMatrix& A)
  {
    A_  = &A;
    op_ = nullptr;
    resize(A.Height());
  }
  void SetOperator(const GradientOperator& op)
  {
    op_ = &op;
    A_  = nullptr;
    resize(op.Height());
  }
  void SetPreconditioner(const JacobiSmoother& prec) { prec_ = &prec; }
  /// mfem::IterativeSolver's default (solvers.cpp:29-30): x holds the initial guess; Newton drivers set it to false on
  /// their linear solver (src/serac/numerics/equation_solver.cpp:96)
  bool iterative_mode = true;

  /// CGSolver::Mult (solvers.cpp:719-897)
  void Mult(const Vector& b, Vector& x) const
  {
    sb_cg_options opt  = opt_;
    opt.iterative_mode = iterative_mode ? 1 : 0;
    const double* dinv = prec_ ? prec_->dinv() : nullptr;
    if (A_) {
      SERAC_B200_CHECK(sb_cg_solve_csr(S_, A_->pattern(), A_->values().Read(), dinv, b.Read(), x.Write(), &opt, &res_, nullptr));
    } else if (op_) {
      SERAC_B200_CHECK(sb_cg_solve_operator(S_, sb_functional_gradient_operator_apply, op_->handle(), dinv, b.Read(),
                                            x.Write(), &opt, &res_, nullptr));
    } else {
      std::fprintf(stderr, "serac::CGSolver::Mult: no operator set\n");
      std::abort();
    }
  }
  bool   GetConverged() const { return res_.converged != 0; }
  int    GetNumIterations() const { return res_.final_iter; }
  double GetInitialNorm() const { return res_.initial_norm; }
  double GetFinalNorm() const { return res_.final_norm; }

 private:
  void resize(int64_t n)
  {
    if (S_ && n == n_) return;
    sb_cg_destroy(S_);
    S_ = nullptr;
    n_ = n;
    SERAC_B200_CHECK(sb_cg_create(n, &S_));
  }
  sb_cg_options           opt_{1e-12, 0.0, 10, 0, 1};  // mfem's IterativeSolver defaults: rel_tol 0 -> set explicitly
  mutable sb_cg_result    res_{};
  sb_cg_t                 S_  = nullptr;
  int64_t                 n_  = -1;
  const DeviceCsrMatrix*  A_  = nullptr;
  const GradientOperator* op_ = nullptr;
  const JacobiSmoother*   prec_ = nullptr;
};

/// mfem::GMRESSolver: serac's default linear solver (solid_mechanics.hpp:50-57)
class GMRESSolver {
 public:
  GMRESSolver() = default;
  GMRESSolver(const GMRESSolver&)            = delete;
  GMRESSolver& operator=(const GMRESSolver&) = delete;
  ~GMRESSolver() { sb_gmres_destroy(G_); }
  void SetRelTol(double t) { opt_.rel_tol = t; }
  void SetAbsTol(double t) { opt_.abs_tol = t; }
  void SetMaxIter(int n) { opt_.max_iter = n; }
  void SetKDim(int m)
  {
    m_ = m;
    n_ = -1;  // reallocate the basis at the next SetOperator
  }
  void SetOperator(const DeviceCsrMatrix& A)
  {
    A_  = &A;
    op_ = nullptr;
    resize(A.Height());
  }
  void SetOperator(const GradientOperator& op)
  {
    op_ = &op;
    A_  = nullptr;
    resize(op.Height());
  }
  void SetPreconditioner(const JacobiSmoother& prec) { prec_ = &prec; }
  /// mfem::IterativeSolver's default (solvers.cpp:29-30): x holds the initial guess; Newton drivers set it to false on
  /// their linear solver (src/serac/numerics/equation_solver.cpp:96)
  bool iterative_mode = true;

  void Mult(const Vector& b, Vector& x) const
  {
    sb_cg_options opt  = opt_;
    opt.iterative_mode = iterative_mode ? 1 : 0;
    const double* dinv = prec_ ? prec_->dinv() : nullptr;
    if (A_) {
      SERAC_B200_CHECK(sb_gmres_solve_csr(G_, A_->pattern(), A_->values().Read(), dinv, b.Read(), x.Write(), &opt, &res_, nullptr));
    } else if (op_) {
      SERAC_B200_CHECK(sb_gmres_solve_operator(G_, sb_functional_gradient_operator_apply, op_->handle(), dinv, b.Read(),
                                               x.Write(), &opt, &res_, nullptr));
    } else {
      std::fprintf(stderr, "serac::GMRESSolver::Mult: no operator set\n");
      std::abort();
    }
  }
  bool   GetConverged() const { return res_.converged != 0; }
  int    GetNumIterations() const { return res_.final_iter; }
  double GetInitialNorm() const { return res_.initial_norm; }
  double GetFinalNorm() const { return res_.final_norm; }

 private:
  void resize(int64_t n)
  {
    if (G_ && n == n_) return;
    sb_gmres_destroy(G_);
    G_ = nullptr;
    n_ = n;
    SERAC_B200_CHECK(sb_gmres_create(n, m_, &G_));
  }
  sb_cg_options           opt_{1e-12, 0.0, 10, 0, 1};
  mutable sb_cg_result    res_{};
  sb_gmres_t              G_  = nullptr;
  int                     m_  = 50;  // mfem's default
  int64_t                 n_  = -1;
  const DeviceCsrMatrix*  A_  = nullptr;
  const GradientOperator* op_ = nullptr;
  const JacobiSmoother*   prec_ = nullptr;
};

}  // namespace serac
