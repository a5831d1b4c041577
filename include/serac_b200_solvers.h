/* serac_b200 C ABI, part 2 -- what sits between Functional and the Newton driver: the
 * linear algebra a Newton step of SolidMechanics / HeatTransfer does with the Functional's outputs, kept on the device so
 * that a step never leaves the GPU: essential-boundary-condition elimination of the assembled Jacobian, its product with
 * a vector, the Jacobi preconditioner and the preconditioned conjugate-gradient iteration -- over the assembled CSR matrix
 * or matrix-free over Gradient::Mult.
 *
 * Reference interfaces replaced (paths relative to the reference checkout):
 *   r.SetSubVector(bcs_.allEssentialTrueDofs(), 0.0)          src/serac/physics/solid_mechanics.hpp:1129, heat_transfer.hpp:719
 *   bcs_.eliminateAllEssentialDofsFromMatrix(*J_)             src/serac/physics/boundary_conditions/boundary_condition_manager.hpp:91-94
 *     = HypreParMatrix::EliminateRowsCols                     mfem/linalg/hypre_parcsr.cpp:476-489 (A -> [A_ii 0; 0 I], Ae = [0 A_ib; A_bi A_bb - I])
 *   mfem::EliminateBC(A, Ae, ess_dofs, X, B)                  mfem/linalg/hypre.cpp:3383-3388 (solid_mechanics.hpp:1756)
 *   HypreParMatrix::Mult (one rank: CSR y = A x)              the operator CGSolver applies
 *   Jacobi preconditioner (HypreSmoother Jacobi / OperatorJacobiSmoother, z = D^-1 r)
 *   mfem::CGSolver::Mult                                      mfem/linalg/solvers.cpp:719-897 (built by serac::buildLinearSolver,
 *                                                             src/serac/numerics/equation_solver.cpp)
 *
 * Conventions as in serac_b200.h: plain pointers and sizes, *_dev = device pointers, int status return (0 = ok, text in
 * sb_last_error()), `stream` = cudaStream_t as void*.  No CPU fallback.  Reductions are deterministic (fixed grid, fixed
 * summation order, no floating-point atomics): the same inputs give the same iterates bit for bit, run to run.
 */
#ifndef SERAC_B200_SOLVERS_H
#define SERAC_B200_SOLVERS_H

#include "serac_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* v[dofs[i]] = value  (mfem::Vector::SetSubVector(const Array<int>&, double); the residual's essential rows,
 * solid_mechanics.hpp:1129) */
int sb_vector_set_subvector(double* v_dev, const int* dofs_dev, int64_t num_dofs, double value, void* stream);
/* v[dofs[i]] = src[dofs[i]]  (the essential entries of the solution vector, solid_mechanics.hpp:1757-1760) */
int sb_vector_copy_subvector(double* v_dev, const double* src_dev, const int* dofs_dev, int64_t num_dofs, void* stream);

/* y += a x  (mfem::Vector::Add(a, x): the Newton update u -= du, solid_mechanics.hpp / equation_solver.cpp) */
int sb_vector_axpy(double* y_dev, double a, const double* x_dev, int64_t n, void* stream);
/* *result_host = (x, y), deterministic (fixed grid, fixed summation order); synchronises `stream` (mfem::Vector::operator*,
 * Norml2 = sqrt of it: the Newton driver's convergence test) */
int sb_vector_dot(const double* x_dev, const double* y_dev, int64_t n, double* result_host, void* stream);

/* y = A x for the CSR triple (pattern object + values): one warp per row, lanes stride over the row's nonzeros
 * (coalesced col_ind / values, x gathered through L2), fixed-order shuffle reduction.  HBM-bound: 12 B per nonzero. */
int sb_csr_spmv(sb_csr_t A, const double* values_dev, const double* x_dev, double* y_dev, void* stream);
/* diag[i] = A(i,i) (0 if the pattern has no diagonal entry in row i) */
int sb_csr_diagonal(sb_csr_t A, const double* values_dev, double* diag_dev, void* stream);

/* l1[i] = sum_j |A(i,j)|: the scaling of hypre's l1-Jacobi smoother (mfem::HypreSmoother::l1Jacobi = serac's
 * Preconditioner::HypreL1Jacobi, HeatTransfer's default, heat_transfer.hpp:34-38; mfem/linalg/hypre.cpp:3630-3632 calls
 * hypre_ParCSRComputeL1Norms(A, 1, ...): "option 1: the l1 norm of the rows", hypre 2.26.0 -- a third-party dependency
 * absent from the reference tree, so this entry restates its published definition: parity unpinned).  One sweep from a
 * zero guess is z = r / l1: pass the result through sb_jacobi_setup and use it as diag_inv_dev. */
int sb_csr_row_l1_norms(sb_csr_t A, const double* values_dev, double* l1_dev, void* stream);

/* HypreParMatrix::EliminateRowsCols at one rank (hypre_parcsr.cpp:476-489): rows and columns `ess_dofs` of A are zeroed
 * and their diagonal entries set to 1, in place, keeping the pattern; values_e_dev (may be NULL) receives, in the SAME
 * pattern, the eliminated part Ae with Ae(b,b) = A(b,b) - 1, so that A_before = A_after + Ae entry by entry. */
int sb_csr_eliminate_rows_cols(sb_csr_t A, double* values_dev, const int* ess_dofs_dev, int64_t num_ess,
                               double* values_e_dev, void* stream);
/* mfem::EliminateBC (hypre.cpp:3383; HypreParMatrix::EliminateBC): B -= Ae X, then B[b] = A(b,b) * X[b] for the
 * essential dofs (A(b,b) = 1 after sb_csr_eliminate_rows_cols).  values_dev: the eliminated matrix. */
int sb_csr_eliminate_bc(sb_csr_t A, const double* values_dev, const double* values_e_dev, const int* ess_dofs_dev,
                        int64_t num_ess, const double* X_dev, double* B_dev, void* stream);

/* an operator y = Op(x) on device vectors: what CGSolver::SetOperator receives */
typedef int (*sb_operator_apply)(void* ctx, const double* x_dev, double* y_dev, void* stream);

/* Matrix-free Jacobian with essential boundary conditions: the operator [J_ii 0; 0 I] applied through
 * sb_functional_jvp (Gradient::Mult, F/functional.hpp:556-559) for the derivative stored by the last evaluate(wrt):
 * y = J (x with essential entries zeroed), then y[b] = x[b].  Equal to sb_csr_spmv of the eliminated assembled matrix up
 * to summation order.  Pass sb_functional_gradient_operator_apply + the handle to sb_cg_solve_operator. */
typedef struct sb_gradient_operator_s* sb_gradient_operator_t;
/* num_dofs: LSize of trial space `wrt` (= of the test space: the operator is square) */
int sb_functional_gradient_operator_create(sb_functional_t F, int wrt, int64_t num_dofs, const int* ess_dofs_dev,
                                           int64_t num_ess, sb_gradient_operator_t* out);
int sb_functional_gradient_operator_destroy(sb_gradient_operator_t op);
int sb_functional_gradient_operator_apply(void* gradient_operator, const double* x_dev, double* y_dev, void* stream);

/* mfem::CGSolver (solvers.cpp:719-897), same iteration, same stopping rule ((B r, r) <= max(rel_tol^2 (B r0, r0),
 * abs_tol^2)), same outputs.  All vectors AND all scalars (nom, den, alpha, beta) stay on the device: an iteration is
 * three launches for a CSR operator -- [x += alpha d, r -= alpha z, z = D^-1 r, (r,z)], [d = z + beta d],
 * [z = A d, (d,z)] -- each reduction finished by the last block to arrive, which also derives alpha / beta and decides
 * convergence into a device-side flag; once the flag is up the remaining enqueued kernels of the solve return at once.
 * The host reads the flag every `check_every` iterations (1 = after every iteration, as the reference does; larger values
 * let the host run ahead; the iterates, final_iter and x do not depend on it). */
typedef struct {
  double rel_tol, abs_tol; /* IterativeSolver::SetRelTol / SetAbsTol */
  int    max_iter;         /* SetMaxIter */
  int    iterative_mode;   /* 1: x holds the initial guess (mfem::IterativeSolver's default, solvers.cpp:29-30);
                              0: x = 0 initially (what Newton drivers set on their linear solver,
                              src/serac/numerics/equation_solver.cpp:96) */
  int    check_every;      /* >= 1 */
} sb_cg_options;

typedef struct {
  int    converged;    /* GetConverged() */
  int    final_iter;   /* GetNumIterations() */
  double initial_norm; /* sqrt((B r0, r0)) */
  double final_norm;   /* GetFinalNorm() = sqrt((B r, r)) at exit */
} sb_cg_result;

typedef struct sb_cg_s* sb_cg_t;
/* work vectors (r, d, z), scalars and reduction partials for systems of `n` unknowns */
int sb_cg_create(int64_t n, sb_cg_t* out);
int sb_cg_destroy(sb_cg_t S);
/* diag_inv_dev: Jacobi preconditioner D^-1 (NULL = no preconditioner).  b_dev, x_dev: n doubles. */
int sb_cg_solve_csr(sb_cg_t S, sb_csr_t A, const double* values_dev, const double* diag_inv_dev, const double* b_dev,
                    double* x_dev, const sb_cg_options* opt, sb_cg_result* result, void* stream);
int sb_cg_solve_operator(sb_cg_t S, sb_operator_apply apply, void* ctx, const double* diag_inv_dev,
                         const double* b_dev, double* x_dev, const sb_cg_options* opt, sb_cg_result* result,
                         void* stream);
/* mfem::GMRESSolver (solvers.cpp:981-1165): restarted GMRES(kdim) with modified Gram-Schmidt and a left Jacobi
 * preconditioner -- serac's default linear solver for SolidMechanics and HeatTransfer (solid_mechanics.hpp:50-57,
 * heat_transfer.hpp:34-38; kdim 50 = mfem's default).  Same iteration, stopping rule (||M r|| <= max(rel_tol ||M r0||,
 * abs_tol)), restart and exits.  The Krylov basis ((kdim + 1) n doubles) and the Hessenberg column under construction
 * live on the device; the host reads one finished column per iteration for the Givens rotations, as the reference does.
 * Options / result: the sb_cg_* structs (check_every is ignored; initial_norm / final_norm are ||M r||). */
typedef struct sb_gmres_s* sb_gmres_t;
int sb_gmres_create(int64_t n, int kdim, sb_gmres_t* out);
int sb_gmres_destroy(sb_gmres_t G);
int sb_gmres_solve_csr(sb_gmres_t G, sb_csr_t A, const double* values_dev, const double* diag_inv_dev,
                       const double* b_dev, double* x_dev, const sb_cg_options* opt, sb_cg_result* result, void* stream);
int sb_gmres_solve_operator(sb_gmres_t G, sb_operator_apply apply, void* ctx, const double* diag_inv_dev,
                            const double* b_dev, double* x_dev, const sb_cg_options* opt, sb_cg_result* result,
                            void* stream);
/* dinv[i] = 1 / diag[i]  (OperatorJacobiSmoother::Setup with damping 1) */
int sb_jacobi_setup(const double* diag_dev, double* diag_inv_dev, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SERAC_B200_SOLVERS_H */
