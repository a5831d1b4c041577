// TEST INFRASTRUCTURE ONLY (oracle/).  The reference stack's own linear algebra for the Newton step, run through the
// serial mfem library that oracle/build_ref_libs.sh compiles from /root/reference/mfem: mfem::CGSolver with
// mfem::OperatorJacobiSmoother on an mfem::SparseMatrix, and essential-BC elimination.  serac does the elimination with
// HypreParMatrix::EliminateRowsCols / mfem::EliminateBC (boundary_condition_manager.hpp:91-94, solid_mechanics.hpp:1756);
// hypre is not in this image, so the elimination here is mfem's serial twin, SparseMatrix::EliminateRowCol(rc, Ae,
// DIAG_ONE) (mfem/linalg/sparsemat.cpp:2284-2380) -- the same A -> [A_ii 0; 0 I], Ae = [0 A_ib; A_bi A_bb - I] split that
// mfem/linalg/hypre_parcsr.cpp:476-489 documents -- and EliminateBC is the three lines of mfem/linalg/hypre.cpp:2422-2478.
// Product code never includes or links this.
#include <vector>

#include "mfem.hpp"

namespace {
/// a SparseMatrix that owns copies of the caller's CSR arrays
mfem::SparseMatrix* make_matrix(int n, const int* row_ptr, const int* col_ind, const double* values)
{
  const int nnz = row_ptr[n];
  int*      I   = mfem::Memory<int>(n + 1);
  int*      J   = mfem::Memory<int>(nnz);
  double*   A   = mfem::Memory<double>(nnz);
  for (int i = 0; i <= n; i++) I[i] = row_ptr[i];
  for (int j = 0; j < nnz; j++) {
    J[j] = col_ind[j];
    A[j] = values[j];
  }
  return new mfem::SparseMatrix(I, J, A, n, n);  // takes ownership
}
}  // namespace

extern "C" {

/// values (in/out): A -> eliminated A; values_e (out, same pattern): Ae
void ref_eliminate_rows_cols(int n, const int* row_ptr, const int* col_ind, double* values, const int* ess, int num_ess,
                             double* values_e)
{
  std::unique_ptr<mfem::SparseMatrix> A(make_matrix(n, row_ptr, col_ind, values));
  mfem::SparseMatrix                  Ae(n, n);
  for (int k = 0; k < num_ess; k++) A->EliminateRowCol(ess[k], Ae, mfem::Operator::DIAG_ONE);
  Ae.Finalize();
  const int nnz = row_ptr[n];
  for (int j = 0; j < nnz; j++) values[j] = A->GetData()[j];
  for (int j = 0; j < nnz; j++) values_e[j] = 0.0;
  const int*    eI = Ae.GetI();
  const int*    eJ = Ae.GetJ();
  const double* eA = Ae.GetData();
  for (int i = 0; i < n; i++)
    for (int k = eI[i]; k < eI[i + 1]; k++)
      for (int j = row_ptr[i]; j < row_ptr[i + 1]; j++)
        if (col_ind[j] == eJ[k]) values_e[j] += eA[k];  // Ae's pattern is a subset of A's
}

/// mfem::EliminateBC (hypre.cpp:2422-2478): b -= Ae x; b(r) = A(r,r) x(r) on the essential rows
void ref_eliminate_bc(int n, const int* row_ptr, const int* col_ind, const double* values, const double* values_e,
                      const int* ess, int num_ess, const double* x, double* b)
{
  std::unique_ptr<mfem::SparseMatrix> A(make_matrix(n, row_ptr, col_ind, values));
  std::unique_ptr<mfem::SparseMatrix> Ae(make_matrix(n, row_ptr, col_ind, values_e));
  mfem::Vector X(const_cast<double*>(x), n), B(b, n);
  Ae->AddMult(X, B, -1.0);
  for (int k = 0; k < num_ess; k++) {
    const int r = ess[k];
    B(r)        = A->Elem(r, r) * X(r);
  }
}

/// mfem::CGSolver::Mult (solvers.cpp:719-897) with an optional OperatorJacobiSmoother built from the matrix diagonal
void ref_cg(int n, const int* row_ptr, const int* col_ind, const double* values, const double* b, double* x, int jacobi,
            double rel_tol, double abs_tol, int max_iter, int iterative_mode, int* converged, int* final_iter,
            double* initial_norm, double* final_norm)
{
  std::unique_ptr<mfem::SparseMatrix> A(make_matrix(n, row_ptr, col_ind, values));
  mfem::Vector                        B(const_cast<double*>(b), n), X(x, n), diag(n);
  A->GetDiag(diag);
  mfem::Array<int>              no_ess;
  mfem::OperatorJacobiSmoother  prec(diag, no_ess, 1.0);
  mfem::CGSolver                cg;
  cg.SetOperator(*A);
  if (jacobi) cg.SetPreconditioner(prec);
  cg.SetRelTol(rel_tol);
  cg.SetAbsTol(abs_tol);
  cg.SetMaxIter(max_iter);
  cg.SetPrintLevel(-1);
  cg.iterative_mode = iterative_mode != 0;
  cg.Mult(B, X);
  *converged    = cg.GetConverged() ? 1 : 0;
  *final_iter   = cg.GetNumIterations();
  *initial_norm = cg.GetInitialNorm();
  *final_norm   = cg.GetFinalNorm();
}

/// mfem::GMRESSolver::Mult (solvers.cpp:981-1165), restart length kdim, optional OperatorJacobiSmoother
void ref_gmres(int n, const int* row_ptr, const int* col_ind, const double* values, const double* b, double* x, int jacobi,
               int kdim, double rel_tol, double abs_tol, int max_iter, int iterative_mode, int* converged, int* final_iter,
               double* initial_norm, double* final_norm)
{
  std::unique_ptr<mfem::SparseMatrix> A(make_matrix(n, row_ptr, col_ind, values));
  mfem::Vector                        B(const_cast<double*>(b), n), X(x, n), diag(n);
  A->GetDiag(diag);
  mfem::Array<int>             no_ess;
  mfem::OperatorJacobiSmoother prec(diag, no_ess, 1.0);
  mfem::GMRESSolver            gmres;
  gmres.SetOperator(*A);
  if (jacobi) gmres.SetPreconditioner(prec);
  gmres.SetKDim(kdim);
  gmres.SetRelTol(rel_tol);
  gmres.SetAbsTol(abs_tol);
  gmres.SetMaxIter(max_iter);
  gmres.SetPrintLevel(-1);
  gmres.iterative_mode = iterative_mode != 0;
  gmres.Mult(B, X);
  *converged    = gmres.GetConverged() ? 1 : 0;
  *final_iter   = gmres.GetNumIterations();
  *initial_norm = gmres.GetInitialNorm();
  *final_norm   = gmres.GetFinalNorm();
}

/// y = A x with mfem::SparseMatrix::Mult
void ref_spmv(int n, const int* row_ptr, const int* col_ind, const double* values, const double* x, double* y)
{
  std::unique_ptr<mfem::SparseMatrix> A(make_matrix(n, row_ptr, col_ind, values));
  mfem::Vector                        X(const_cast<double*>(x), n), Y(y, n);
  A->Mult(X, Y);
}
}
