// serac_b200: the linear algebra between Functional and the Newton driver (include/serac_b200_solvers.h), on the device:
// essential-BC elimination of the assembled Jacobian, CSR y = A x, Jacobi, and mfem::CGSolver's iteration with every
// vector and scalar resident in HBM.  sm_100a only; no CPU fallback.
//
// Every kernel here is HBM-bound streaming work (SpMV: 12 B per nonzero; the vector updates: 24-40 B per unknown), so the
// rules that matter are the ones about coalescing and grid sizing: one warp per CSR row with the lanes striding over the
// row's nonzeros, grid-stride vector kernels on a fixed grid of SM-count multiples, no floating-point atomics.  A dot
// product is finished by the last block to arrive (ticket counter), which sums the per-block partials in index order --
// the order of arrival never enters the sum -- and derives the CG scalars on the spot, so that no reduction needs a
// launch of its own or a round trip to the host.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <memory>
#include <vector>

#include "common.hpp"
#include "serac_b200_solvers.h"

namespace {

constexpr int kThreads      = 256;
constexpr int kWarpsPerCta  = kThreads / 32;
constexpr int kMaxPartials  = 148 * 8;  // fixed upper bound of every reducing grid: 8 CTAs of 256 threads per SM

inline cudaStream_t as_stream(void* s) { return cudaStream_t(s); }

int sm_count()
{
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        n <= 0)
      n = 148;
  }
  return n;
}

/// grid of a grid-stride kernel over `work` items handled `per_cta` at a time: whole multiples of the SM count, capped
int reducing_grid(int64_t work, int per_cta)
{
  const int64_t want = (work + per_cta - 1) / per_cta;
  const int64_t cap  = std::min<int64_t>(int64_t(sm_count()) * 8, kMaxPartials);
  return int(std::max<int64_t>(1, std::min<int64_t>(want, cap)));
}

// ------------------------------------------------------------------------------------------------------------------
// scalars of one solve, resident on the device
// ------------------------------------------------------------------------------------------------------------------
struct CgScalars {
  double nom, nom0, den, alpha, beta, betanom, r0;
  double rel_tol, abs_tol;
  double initial_norm, final_norm;
  int    iter;       // the iteration whose update kernel ran last (mfem's loop variable i)
  int    max_iter;
  int    done;       // 1: the solve has ended; every later kernel of this solve returns at once
  int    converged;
  int    final_iter;
  int    pad;
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;  // lane 0 holds the sum; fixed tree
}

/// block-wide sum in a fixed order; the result is valid in thread 0
__device__ __forceinline__ double block_sum(double v, double* smem /* kWarpsPerCta */)
{
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();  // smem may still be read from a previous use
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < kWarpsPerCta; w++) s += smem[w];
  }
  return s;
}

/// Publishes this block's partial and tells whether this block is the last one of the grid to do so.  When it is, all
/// partials are visible to it (fence + ticket), and it returns their sum (in thread 0), taken in index order.
__device__ __forceinline__ bool grid_sum_last_block(double block_partial, double* __restrict__ partials,
                                                    unsigned int* __restrict__ ticket, double* smem, double* total)
{
  __shared__ bool last;
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = block_partial;
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    last                 = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return false;
  __threadfence();
  double v = 0.0;
  for (int i = threadIdx.x; i < int(gridDim.x); i += blockDim.x) v += __ldcg(partials + i);
  const double s = block_sum(v, smem);
  if (threadIdx.x == 0) {
    *total  = s;
    *ticket = 0u;  // ready for the next reduction on this stream
  }
  return true;
}

// ------------------------------------------------------------------------------------------------------------------
// small vector kernels
// ------------------------------------------------------------------------------------------------------------------
__global__ void set_subvector_kernel(double* __restrict__ v, const int* __restrict__ dofs, int64_t n, double value)
{
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    v[dofs[i]] = value;
}

__global__ void copy_subvector_kernel(double* __restrict__ v, const double* __restrict__ src, const int* __restrict__ dofs,
                                      int64_t n)
{
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const int k = dofs[i];
    v[k]        = src[k];
  }
}

__global__ void mark_kernel(unsigned char* __restrict__ mark, const int* __restrict__ dofs, int64_t n)
{
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    mark[dofs[i]] = 1;
}

__global__ void reciprocal_kernel(const double* __restrict__ d, double* __restrict__ dinv, int64_t n)
{
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    dinv[i] = 1.0 / d[i];
}

__global__ void axpy_kernel(double* __restrict__ y, double a, const double* __restrict__ x, int64_t n)
{
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    y[i] = y[i] + a * x[i];
}

/// out[0] = (x, y); out[1] is the ticket counter (zero before and after)
__global__ void __launch_bounds__(kThreads) dot_kernel(const double* __restrict__ x, const double* __restrict__ y, int64_t n,
                                                       double* __restrict__ out, double* __restrict__ partials,
                                                       unsigned int* __restrict__ ticket)
{
  __shared__ double smem[kWarpsPerCta];
  double            acc = 0.0;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    acc += x[i] * y[i];
  const double part  = block_sum(acc, smem);
  double       total = 0.0;
  if (grid_sum_last_block(part, partials, ticket, smem, &total) && threadIdx.x == 0) *out = total;
}

/// y = x with the marked entries zeroed
__global__ void masked_copy_kernel(const double* __restrict__ x, const unsigned char* __restrict__ mark,
                                   double* __restrict__ y, int64_t n)
{
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    y[i] = mark[i] ? 0.0 : x[i];
}

// ------------------------------------------------------------------------------------------------------------------
// CSR kernels: one warp per row
// ------------------------------------------------------------------------------------------------------------------
/// the row sum of A(row, :) x in a fixed order: lane l takes nonzeros l, l + 32, ... then the shuffle tree
__device__ __forceinline__ double row_dot(const int* __restrict__ col_ind, const double* __restrict__ values,
                                          const double* __restrict__ x, int begin, int end, int lane)
{
  double acc = 0.0;
  for (int j = begin + lane; j < end; j += 32) acc += values[j] * __ldg(x + col_ind[j]);
  return warp_sum(acc);
}

__global__ void __launch_bounds__(kThreads) csr_spmv_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col_ind,
                                                            const double* __restrict__ values,
                                                            const double* __restrict__ x, double* __restrict__ y,
                                                            int64_t num_rows)
{
  const int     lane  = threadIdx.x & 31;
  const int64_t warp0 = blockIdx.x * int64_t(kWarpsPerCta) + (threadIdx.x >> 5);
  for (int64_t row = warp0; row < num_rows; row += int64_t(gridDim.x) * kWarpsPerCta) {
    const double s = row_dot(col_ind, values, x, row_ptr[row], row_ptr[row + 1], lane);
    if (lane == 0) y[row] = s;
  }
}

__global__ void __launch_bounds__(kThreads) csr_diagonal_kernel(const int* __restrict__ row_ptr,
                                                                const int* __restrict__ col_ind,
                                                                const double* __restrict__ values,
                                                                double* __restrict__ diag, int64_t num_rows)
{
  const int     lane  = threadIdx.x & 31;
  const int64_t warp0 = blockIdx.x * int64_t(kWarpsPerCta) + (threadIdx.x >> 5);
  for (int64_t row = warp0; row < num_rows; row += int64_t(gridDim.x) * kWarpsPerCta) {
    double d = 0.0;  // at most one lane finds the diagonal entry (col_ind is unique within a row)
    for (int j = row_ptr[row] + lane; j < row_ptr[row + 1]; j += 32)
      if (col_ind[j] == int(row)) d = values[j];
    d = warp_sum(d);
    if (lane == 0) diag[row] = d;
  }
}

/// l1[row] = sum_j |A(row, j)|, same lane / tree order as the row products
__global__ void __launch_bounds__(kThreads) csr_row_l1_kernel(const int* __restrict__ row_ptr,
                                                              const double* __restrict__ values, double* __restrict__ l1,
                                                              int64_t num_rows)
{
  const int     lane  = threadIdx.x & 31;
  const int64_t warp0 = blockIdx.x * int64_t(kWarpsPerCta) + (threadIdx.x >> 5);
  for (int64_t row = warp0; row < num_rows; row += int64_t(gridDim.x) * kWarpsPerCta) {
    double acc = 0.0;
    for (int j = row_ptr[row] + lane; j < row_ptr[row + 1]; j += 32) acc += fabs(values[j]);
    acc = warp_sum(acc);
    if (lane == 0) l1[row] = acc;
  }
}

/// A -> [A_ii 0; 0 I], Ae = [0 A_ib; A_bi A_bb - I]   (hypre_parcsr.cpp:476-489)
__global__ void __launch_bounds__(kThreads) csr_eliminate_kernel(const int* __restrict__ row_ptr,
                                                                 const int* __restrict__ col_ind,
                                                                 double* __restrict__ values, double* __restrict__ values_e,
                                                                 const unsigned char* __restrict__ mark, int64_t num_rows)
{
  const int     lane  = threadIdx.x & 31;
  const int64_t warp0 = blockIdx.x * int64_t(kWarpsPerCta) + (threadIdx.x >> 5);
  for (int64_t row = warp0; row < num_rows; row += int64_t(gridDim.x) * kWarpsPerCta) {
    const bool row_ess = mark[row] != 0;
    for (int j = row_ptr[row] + lane; j < row_ptr[row + 1]; j += 32) {
      const int  col  = col_ind[j];
      const bool elim = row_ess || mark[col] != 0;
      double     a = values[j], ae = 0.0;
      if (elim) {
        const double keep = (col == int(row)) ? 1.0 : 0.0;  // an eliminated diagonal entry becomes 1
        ae                = a - keep;
        a                 = keep;
        values[j]         = a;
      }
      if (values_e) values_e[j] = ae;
    }
  }
}

/// B -= Ae X on every row (rows of Ae without eliminated entries hold zeros), warp per row
__global__ void __launch_bounds__(kThreads) csr_eliminate_bc_rows_kernel(const int* __restrict__ row_ptr,
                                                                         const int* __restrict__ col_ind,
                                                                         const double* __restrict__ values_e,
                                                                         const double* __restrict__ X, double* __restrict__ B,
                                                                         int64_t num_rows)
{
  const int     lane  = threadIdx.x & 31;
  const int64_t warp0 = blockIdx.x * int64_t(kWarpsPerCta) + (threadIdx.x >> 5);
  for (int64_t row = warp0; row < num_rows; row += int64_t(gridDim.x) * kWarpsPerCta) {
    const double s = row_dot(col_ind, values_e, X, row_ptr[row], row_ptr[row + 1], lane);
    if (lane == 0) B[row] -= s;
  }
}

/// B[b] = A(b,b) X[b] for the essential dofs: one warp per essential row finds its diagonal entry
__global__ void __launch_bounds__(kThreads) csr_eliminate_bc_ess_kernel(const int* __restrict__ row_ptr,
                                                                        const int* __restrict__ col_ind,
                                                                        const double* __restrict__ values,
                                                                        const int* __restrict__ ess, int64_t num_ess,
                                                                        const double* __restrict__ X, double* __restrict__ B)
{
  const int     lane  = threadIdx.x & 31;
  const int64_t warp0 = blockIdx.x * int64_t(kWarpsPerCta) + (threadIdx.x >> 5);
  for (int64_t k = warp0; k < num_ess; k += int64_t(gridDim.x) * kWarpsPerCta) {
    const int row = ess[k];
    double    d   = 0.0;
    for (int j = row_ptr[row] + lane; j < row_ptr[row + 1]; j += 32)
      if (col_ind[j] == row) d = values[j];
    d = warp_sum(d);
    if (lane == 0) B[row] = d * X[row];
  }
}

// ------------------------------------------------------------------------------------------------------------------
// CG kernels (mfem/linalg/solvers.cpp:719-897; line numbers in the comments refer to that file)
// ------------------------------------------------------------------------------------------------------------------
__global__ void cg_reset_kernel(CgScalars* S, double rel_tol, double abs_tol, int max_iter, unsigned int* ticket)
{
  CgScalars z{};
  z.rel_tol  = rel_tol;
  z.abs_tol  = abs_tol;
  z.max_iter = max_iter;
  *S         = z;
  *ticket    = 0u;
}

/// r = b - r (r holds A x on entry) or r = b, x = 0  (:725-734)
__global__ void cg_initial_residual_kernel(const double* __restrict__ b, double* __restrict__ r, double* __restrict__ x,
                                           int iterative_mode, int64_t n)
{
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    if (iterative_mode) {
      r[i] = b[i] - r[i];
    } else {
      r[i] = b[i];
      x[i] = 0.0;
    }
  }
}

/// z = B r, d = z (or d = r), nom0 = nom = (d, r); the early exits of :755-775
__global__ void __launch_bounds__(kThreads) cg_start_kernel(const double* __restrict__ r, const double* __restrict__ dinv,
                                                            double* __restrict__ z, double* __restrict__ d, int64_t n,
                                                            CgScalars* __restrict__ S, double* __restrict__ partials,
                                                            unsigned int* __restrict__ ticket)
{
  __shared__ double smem[kWarpsPerCta];
  double            acc = 0.0;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const double ri = r[i];
    const double di = dinv ? dinv[i] * ri : ri;
    if (dinv) z[i] = di;
    d[i] = di;
    acc += di * ri;
  }
  const double part = block_sum(acc, smem);
  double       nom  = 0.0;
  if (grid_sum_last_block(part, partials, ticket, smem, &nom) && threadIdx.x == 0) {
    S->nom = S->nom0 = nom;
    S->betanom       = nom;
    if (nom >= 0.0) S->initial_norm = sqrt(nom);
    if (!(nom >= 0.0)) {  // the preconditioner is not positive definite (:755-767); NaN ends the solve too (:747)
      S->done         = 1;
      S->converged    = 0;
      S->final_iter   = 0;
      S->initial_norm = nom;
      S->final_norm   = nom;
    } else {
      S->r0 = fmax(nom * S->rel_tol * S->rel_tol, S->abs_tol * S->abs_tol);
      if (nom <= S->r0) {  // :769-775
        S->done       = 1;
        S->converged  = 1;
        S->final_iter = 0;
        S->final_norm = sqrt(nom);
      }
    }
    __threadfence();
  }
}

/// what follows den = (d, A d): :778-794 before the loop, :857-872 inside it
__device__ __forceinline__ void cg_after_den(CgScalars* S, double den)
{
  S->den = den;
  if (S->iter > 0) {      // inside the loop: this product belongs to iteration iter + 1
    S->nom = S->betanom;  // :872 -- harmless when den == 0 ends the solve below
  }
  if (den == 0.0 || !isfinite(den)) {
    S->done       = 1;
    S->converged  = 0;
    S->final_iter = (S->iter > 0) ? S->iter + 1 : 0;
    S->final_norm = sqrt(S->betanom);
    return;
  }
  S->alpha = S->nom / den;  // :801
}

/// z = A d and den = (d, z) for a CSR operator, one warp per row
__global__ void __launch_bounds__(kThreads) cg_spmv_dot_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col_ind,
                                                               const double* __restrict__ values,
                                                               const double* __restrict__ d, double* __restrict__ z,
                                                               int64_t num_rows, CgScalars* __restrict__ S,
                                                               double* __restrict__ partials,
                                                               unsigned int* __restrict__ ticket)
{
  if (*(volatile int*)&S->done) return;
  __shared__ double smem[kWarpsPerCta];
  const int         lane  = threadIdx.x & 31;
  const int64_t     warp0 = blockIdx.x * int64_t(kWarpsPerCta) + (threadIdx.x >> 5);
  double            acc   = 0.0;  // lane 0 of each warp accumulates d[row] * z[row] over the warp's rows, ascending
  for (int64_t row = warp0; row < num_rows; row += int64_t(gridDim.x) * kWarpsPerCta) {
    const double s = row_dot(col_ind, values, d, row_ptr[row], row_ptr[row + 1], lane);
    if (lane == 0) {
      z[row] = s;
      acc += d[row] * s;
    }
  }
  const double part = block_sum(acc, smem);
  double       den  = 0.0;
  if (grid_sum_last_block(part, partials, ticket, smem, &den) && threadIdx.x == 0) {
    cg_after_den(S, den);
    __threadfence();
  }
}

/// den = (d, z) for an operator applied by its own kernels
__global__ void __launch_bounds__(kThreads) cg_dot_kernel(const double* __restrict__ d, const double* __restrict__ z, int64_t n,
                                                          CgScalars* __restrict__ S, double* __restrict__ partials,
                                                          unsigned int* __restrict__ ticket)
{
  if (*(volatile int*)&S->done) return;
  __shared__ double smem[kWarpsPerCta];
  double            acc = 0.0;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    acc += d[i] * z[i];
  const double part = block_sum(acc, smem);
  double       den  = 0.0;
  if (grid_sum_last_block(part, partials, ticket, smem, &den) && threadIdx.x == 0) {
    cg_after_den(S, den);
    __threadfence();
  }
}

/// x += alpha d, r -= alpha z, z = B r, betanom = (r, z) (or (r, r)); the exits of :815-845 and beta (:847)
__global__ void __launch_bounds__(kThreads) cg_update_kernel(double* __restrict__ x, double* __restrict__ r, double* __restrict__ z,
                                                             const double* __restrict__ d, const double* __restrict__ dinv,
                                                             int64_t n, CgScalars* __restrict__ S,
                                                             double* __restrict__ partials,
                                                             unsigned int* __restrict__ ticket)
{
  if (*(volatile int*)&S->done) return;
  __shared__ double smem[kWarpsPerCta];
  const double      alpha = S->alpha;
  double            acc   = 0.0;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const double di = d[i];
    x[i]            = x[i] + alpha * di;
    const double ri = r[i] + (-alpha) * z[i];  // add(r, -alpha, z, r)
    r[i]            = ri;
    if (dinv) {
      const double zi = dinv[i] * ri;
      z[i]            = zi;
      acc += ri * zi;
    } else {
      acc += ri * ri;
    }
  }
  const double part    = block_sum(acc, smem);
  double       betanom = 0.0;
  if (grid_sum_last_block(part, partials, ticket, smem, &betanom) && threadIdx.x == 0) {
    const int i = S->iter + 1;
    S->iter     = i;
    S->betanom  = betanom;
    if (!(betanom >= 0.0)) {  // :815-825 (and NaN, :814)
      S->done       = 1;
      S->converged  = 0;
      S->final_iter = i;
      S->final_norm = sqrt(betanom);
    } else if (betanom <= S->r0) {  // :835-840
      S->done       = 1;
      S->converged  = 1;
      S->final_iter = i;
      S->final_norm = sqrt(betanom);
    } else if (i + 1 > S->max_iter) {  // :842-845, final_iter = max_iter (:798)
      S->done       = 1;
      S->converged  = 0;
      S->final_iter = S->max_iter;
      S->final_norm = sqrt(betanom);
    } else {
      S->beta = betanom / S->nom;  // :847
    }
    __threadfence();
  }
}

/// d = z + beta d (or r + beta d)  (:848-855)
__global__ void cg_direction_kernel(double* __restrict__ d, const double* __restrict__ zr, int64_t n,
                                    const CgScalars* __restrict__ S)
{
  if (*(volatile const int*)&S->done) return;
  const double beta = S->beta;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    d[i] = zr[i] + beta * d[i];
}

struct CsrView {
  const int *row_ptr = nullptr, *col_ind = nullptr;
  int64_t    num_rows = 0, nnz = 0;
};

int view_of(sb_csr_t A, CsrView* v)
{
  SB_REQUIRE(A != nullptr, "null CSR pattern");
  if (int e = sb_csr_pattern_dev(A, &v->row_ptr, &v->col_ind)) return e;
  v->num_rows = sb_csr_num_rows(A);
  v->nnz      = sb_csr_nnz(A);
  return SB_OK;
}

template <class T>
int dev_new(T** p, size_t count)
{
  SB_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(count, 1) * sizeof(T)));
  return SB_OK;
}

int build_mark(const int* ess_dofs_dev, int64_t num_ess, int64_t n, cudaStream_t s, unsigned char** mark)
{
  if (int e = dev_new(mark, size_t(n))) return e;
  SB_CHECK_CUDA(cudaMemsetAsync(*mark, 0, size_t(std::max<int64_t>(n, 1)), s));
  if (num_ess > 0) {
    mark_kernel<<<reducing_grid(num_ess, kThreads), kThreads, 0, s>>>(*mark, ess_dofs_dev, num_ess);
    sb_count_launch();
    SB_CHECK_CUDA(cudaGetLastError());
  }
  return SB_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------------
extern "C" int sb_vector_set_subvector(double* v_dev, const int* dofs_dev, int64_t num_dofs, double value, void* stream)
{
  SB_REQUIRE(num_dofs >= 0, "sb_vector_set_subvector: negative count");
  if (num_dofs == 0) return SB_OK;
  SB_REQUIRE(v_dev && dofs_dev, "sb_vector_set_subvector: null pointer");
  set_subvector_kernel<<<reducing_grid(num_dofs, kThreads), kThreads, 0, as_stream(stream)>>>(v_dev, dofs_dev, num_dofs,
                                                                                             value);
  sb_count_launch();
  SB_CHECK_CUDA(cudaGetLastError());
  return SB_OK;
}

extern "C" int sb_vector_copy_subvector(double* v_dev, const double* src_dev, const int* dofs_dev, int64_t num_dofs,
                                        void* stream)
{
  SB_REQUIRE(num_dofs >= 0, "sb_vector_copy_subvector: negative count");
  if (num_dofs == 0) return SB_OK;
  SB_REQUIRE(v_dev && src_dev && dofs_dev, "sb_vector_copy_subvector: null pointer");
  copy_subvector_kernel<<<reducing_grid(num_dofs, kThreads), kThreads, 0, as_stream(stream)>>>(v_dev, src_dev, dofs_dev,
                                                                                              num_dofs);
  sb_count_launch();
  SB_CHECK_CUDA(cudaGetLastError());
  return SB_OK;
}

extern "C" int sb_vector_axpy(double* y_dev, double a, const double* x_dev, int64_t n, void* stream)
{
  SB_REQUIRE(n >= 0, "sb_vector_axpy: negative size");
  if (n == 0) return SB_OK;
  SB_REQUIRE(y_dev && x_dev, "sb_vector_axpy: null pointer");
  axpy_kernel<<<reducing_grid(n, kThreads), kThreads, 0, as_stream(stream)>>>(y_dev, a, x_dev, n);
  sb_count_launch();
  SB_CHECK_CUDA(cudaGetLastError());
  return SB_OK;
}

extern "C" int sb_vector_dot(const double* x_dev, const double* y_dev, int64_t n, double* result_host, void* stream)
{
  SB_REQUIRE(n >= 0 && result_host, "sb_vector_dot: bad argument");
  *result_host = 0.0;
  if (n == 0) return SB_OK;
  SB_REQUIRE(x_dev && y_dev, "sb_vector_dot: null pointer");
  cudaStream_t s = as_stream(stream);
  // scratch of this call: [0] result, [1] ticket (as 8 bytes), [2 ...] per-block partials
  double* scratch = nullptr;
  if (int e = dev_new(&scratch, size_t(kMaxPartials) + 2)) return e;
  int rc = SB_OK;
  if (cudaMemsetAsync(scratch, 0, 2 * sizeof(double), s) != cudaSuccess) rc = SB_ERR_CUDA;
  if (!rc) {
    dot_kernel<<<reducing_grid(n, kThreads), kThreads, 0, s>>>(x_dev, y_dev, n, scratch, scratch + 2,
                                                               reinterpret_cast<unsigned int*>(scratch + 1));
    sb_count_launch();
    if (cudaGetLastError() != cudaSuccess ||
        cudaMemcpyAsync(result_host, scratch, sizeof(double), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess)
      rc = SB_ERR_CUDA;
  }
  if (rc) sb::set_error("sb_vector_dot: CUDA failure");
  cudaFree(scratch);
  return rc;
}

extern "C" int sb_jacobi_setup(const double* diag_dev, double* diag_inv_dev, int64_t n, void* stream)
{
  SB_REQUIRE(n >= 0, "sb_jacobi_setup: negative size");
  if (n == 0) return SB_OK;
  SB_REQUIRE(diag_dev && diag_inv_dev, "sb_jacobi_setup: null pointer");
  reciprocal_kernel<<<reducing_grid(n, kThreads), kThreads, 0, as_stream(stream)>>>(diag_dev, diag_inv_dev, n);
  sb_count_launch();
  SB_CHECK_CUDA(cudaGetLastError());
  return SB_OK;
}

extern "C" int sb_csr_spmv(sb_csr_t A, const double* values_dev, const double* x_dev, double* y_dev, void* stream)
{
  CsrView v;
  if (int e = view_of(A, &v)) return e;
  if (v.num_rows == 0) return SB_OK;
  SB_REQUIRE(values_dev && x_dev && y_dev, "sb_csr_spmv: null pointer");
  csr_spmv_kernel<<<reducing_grid(v.num_rows, kWarpsPerCta), kThreads, 0, as_stream(stream)>>>(
      v.row_ptr, v.col_ind, values_dev, x_dev, y_dev, v.num_rows);
  sb_count_launch();
  SB_CHECK_CUDA(cudaGetLastError());
  return SB_OK;
}

extern "C" int sb_csr_diagonal(sb_csr_t A, const double* values_dev, double* diag_dev, void* stream)
{
  CsrView v;
  if (int e = view_of(A, &v)) return e;
  if (v.num_rows == 0) return SB_OK;
  SB_REQUIRE(values_dev && diag_dev, "sb_csr_diagonal: null pointer");
  csr_diagonal_kernel<<<reducing_grid(v.num_rows, kWarpsPerCta), kThreads, 0, as_stream(stream)>>>(
      v.row_ptr, v.col_ind, values_dev, diag_dev, v.num_rows);
  sb_count_launch();
  SB_CHECK_CUDA(cudaGetLastError());
  return SB_OK;
}

extern "C" int sb_csr_row_l1_norms(sb_csr_t A, const double* values_dev, double* l1_dev, void* stream)
{
  CsrView v;
  if (int e = view_of(A, &v)) return e;
  if (v.num_rows == 0) return SB_OK;
  SB_REQUIRE(values_dev && l1_dev, "sb_csr_row_l1_norms: null pointer");
  csr_row_l1_kernel<<<reducing_grid(v.num_rows, kWarpsPerCta), kThreads, 0, as_stream(stream)>>>(v.row_ptr, values_dev,
                                                                                                  l1_dev, v.num_rows);
  sb_count_launch();
  SB_CHECK_CUDA(cudaGetLastError());
  return SB_OK;
}

extern "C" int sb_csr_eliminate_rows_cols(sb_csr_t A, double* values_dev, const int* ess_dofs_dev, int64_t num_ess,
                                          double* values_e_dev, void* stream)
{
  CsrView v;
  if (int e = view_of(A, &v)) return e;
  SB_REQUIRE(num_ess >= 0 && (num_ess == 0 || ess_dofs_dev), "sb_csr_eliminate_rows_cols: bad essential dof list");
  SB_REQUIRE(v.num_rows == 0 || values_dev, "sb_csr_eliminate_rows_cols: null values");
  if (v.num_rows == 0) return SB_OK;
  // rows AND columns are eliminated, so the matrix must be square -- as for mfem's EliminateRowsCols
  // (mfem/linalg/hypre_parcsr.cpp:476-489); the marker array is indexed by column as well as by row
  SB_REQUIRE(sb_csr_num_cols(A) == v.num_rows, "sb_csr_eliminate_rows_cols: the matrix is not square");
  cudaStream_t   s    = as_stream(stream);
  // scratch of this call from the stream-ordered allocator: no device synchronisation, and the whole Newton step stays
  // enqueued on the caller's stream
  unsigned char* mark = nullptr;
  SB_CHECK_CUDA(cudaMallocAsync((void**)&mark, size_t(v.num_rows), s));
  cudaError_t err = cudaMemsetAsync(mark, 0, size_t(v.num_rows), s);
  if (err == cudaSuccess && num_ess > 0) {
    mark_kernel<<<reducing_grid(num_ess, kThreads), kThreads, 0, s>>>(mark, ess_dofs_dev, num_ess);
    sb_count_launch();
    err = cudaGetLastError();
  }
  if (err == cudaSuccess) {
    csr_eliminate_kernel<<<reducing_grid(v.num_rows, kWarpsPerCta), kThreads, 0, s>>>(v.row_ptr, v.col_ind, values_dev,
                                                                                      values_e_dev, mark, v.num_rows);
    sb_count_launch();
    err = cudaGetLastError();
  }
  cudaFreeAsync(mark, s);
  SB_CHECK_CUDA(err);
  return SB_OK;
}

extern "C" int sb_csr_eliminate_bc(sb_csr_t A, const double* values_dev, const double* values_e_dev,
                                   const int* ess_dofs_dev, int64_t num_ess, const double* X_dev, double* B_dev,
                                   void* stream)
{
  CsrView v;
  if (int e = view_of(A, &v)) return e;
  SB_REQUIRE(num_ess >= 0 && (num_ess == 0 || ess_dofs_dev), "sb_csr_eliminate_bc: bad essential dof list");
  if (v.num_rows == 0) return SB_OK;
  SB_REQUIRE(values_dev && values_e_dev && X_dev && B_dev, "sb_csr_eliminate_bc: null pointer");
  cudaStream_t s = as_stream(stream);
  csr_eliminate_bc_rows_kernel<<<reducing_grid(v.num_rows, kWarpsPerCta), kThreads, 0, s>>>(v.row_ptr, v.col_ind,
                                                                                            values_e_dev, X_dev, B_dev,
                                                                                            v.num_rows);
  sb_count_launch();
  SB_CHECK_CUDA(cudaGetLastError());
  if (num_ess > 0) {
    csr_eliminate_bc_ess_kernel<<<reducing_grid(num_ess, kWarpsPerCta), kThreads, 0, s>>>(
        v.row_ptr, v.col_ind, values_dev, ess_dofs_dev, num_ess, X_dev, B_dev);
    sb_count_launch();
    SB_CHECK_CUDA(cudaGetLastError());
  }
  return SB_OK;
}

// ---- matrix-free Jacobian with essential boundary conditions --------------------------------------------------------
struct sb_gradient_operator_s {
  sb_functional_t F       = nullptr;
  int             wrt     = 0;
  int64_t         n       = 0;
  int64_t         num_ess = 0;
  int*            ess     = nullptr;  // own copy of the list
  unsigned char*  mark    = nullptr;
  double*         masked  = nullptr;  // x with the essential entries zeroed
};

extern "C" int sb_functional_gradient_operator_destroy(sb_gradient_operator_t op)
{
  if (!op) return SB_OK;
  cudaFree(op->ess);
  cudaFree(op->mark);
  cudaFree(op->masked);
  delete op;
  return SB_OK;
}

extern "C" int sb_functional_gradient_operator_create(sb_functional_t F, int wrt, int64_t num_dofs, const int* ess_dofs_dev,
                                                      int64_t num_ess, sb_gradient_operator_t* out)
{
  SB_REQUIRE(F && out && wrt >= 0 && num_dofs >= 0, "sb_functional_gradient_operator_create: bad argument");
  SB_REQUIRE(num_ess >= 0 && (num_ess == 0 || ess_dofs_dev), "sb_functional_gradient_operator_create: bad dof list");
  auto op     = std::make_unique<sb_gradient_operator_s>();
  op->F       = F;
  op->wrt     = wrt;
  op->n       = num_dofs;
  op->num_ess = num_ess;
  int rc      = SB_OK;
  if (num_ess > 0) {
    rc = dev_new(&op->ess, size_t(num_ess));
    if (!rc && cudaMemcpy(op->ess, ess_dofs_dev, size_t(num_ess) * sizeof(int), cudaMemcpyDeviceToDevice) != cudaSuccess) {
      sb::set_error("sb_functional_gradient_operator_create: copy of the essential dof list failed");
      rc = SB_ERR_CUDA;
    }
    if (!rc) rc = build_mark(op->ess, num_ess, num_dofs, nullptr, &op->mark);
    if (!rc) rc = dev_new(&op->masked, size_t(num_dofs));
    if (!rc && cudaStreamSynchronize(nullptr) != cudaSuccess) rc = SB_ERR_CUDA;
  }
  if (rc) {
    sb_functional_gradient_operator_destroy(op.release());
    return rc;
  }
  *out = op.release();
  return SB_OK;
}

extern "C" int sb_functional_gradient_operator_apply(void* gradient_operator, const double* x_dev, double* y_dev,
                                                     void* stream)
{
  auto* op = static_cast<sb_gradient_operator_t>(gradient_operator);
  SB_REQUIRE(op && x_dev && y_dev, "sb_functional_gradient_operator_apply: null pointer");
  cudaStream_t  s  = as_stream(stream);
  const double* in = x_dev;
  if (op->num_ess > 0) {
    masked_copy_kernel<<<reducing_grid(op->n, kThreads), kThreads, 0, s>>>(x_dev, op->mark, op->masked, op->n);
    sb_count_launch();
    SB_CHECK_CUDA(cudaGetLastError());
    in = op->masked;
  }
  if (int e = sb_functional_jvp(op->F, op->wrt, in, SB_DEVICE, y_dev, SB_DEVICE, stream)) return e;
  if (op->num_ess > 0) return sb_vector_copy_subvector(y_dev, x_dev, op->ess, op->num_ess, stream);
  return SB_OK;
}

// ---- conjugate gradients ----------------------------------------------------------------------------------------------
struct sb_cg_s {
  int64_t       n = 0;
  double *      r = nullptr, *d = nullptr, *z = nullptr;
  double*       partials = nullptr;
  unsigned int* ticket   = nullptr;
  CgScalars*    scalars  = nullptr;
  CgScalars*    host     = nullptr;  // pinned mirror the host polls
};

extern "C" int sb_cg_destroy(sb_cg_t S)
{
  if (!S) return SB_OK;
  cudaFree(S->r);
  cudaFree(S->d);
  cudaFree(S->z);
  cudaFree(S->partials);
  cudaFree(S->ticket);
  cudaFree(S->scalars);
  if (S->host) cudaFreeHost(S->host);
  delete S;
  return SB_OK;
}

extern "C" int sb_cg_create(int64_t n, sb_cg_t* out)
{
  SB_REQUIRE(out && n >= 0, "sb_cg_create: bad argument");
  auto S = std::make_unique<sb_cg_s>();
  S->n   = n;
  int rc = dev_new(&S->r, size_t(n));
  if (!rc) rc = dev_new(&S->d, size_t(n));
  if (!rc) rc = dev_new(&S->z, size_t(n));
  if (!rc) rc = dev_new(&S->partials, size_t(kMaxPartials));
  if (!rc) rc = dev_new(&S->ticket, 1);
  if (!rc) rc = dev_new(&S->scalars, 1);
  if (!rc && cudaMallocHost(reinterpret_cast<void**>(&S->host), sizeof(CgScalars)) != cudaSuccess) {
    sb::set_error("sb_cg_create: cudaMallocHost failed");
    rc = SB_ERR_CUDA;
  }
  if (rc) {
    sb_cg_destroy(S.release());
    return rc;
  }
  *out = S.release();
  return SB_OK;
}

namespace {

struct CgOperator {
  // CSR
  CsrView       A;
  const double* values = nullptr;
  // or a callback
  sb_operator_apply apply = nullptr;
  void*             ctx   = nullptr;
};

/// z = Op(d) [, den = (d, z) into the scalars when `with_dot`]
int apply_operator(sb_cg_t S, const CgOperator& op, const double* d, double* z, bool with_dot, cudaStream_t s)
{
  const int64_t n = S->n;
  if (op.apply) {
    if (int e = op.apply(op.ctx, d, z, s)) return e;
    if (with_dot) {
      cg_dot_kernel<<<reducing_grid(n, kThreads), kThreads, 0, s>>>(d, z, n, S->scalars, S->partials, S->ticket);
      sb_count_launch();
      SB_CHECK_CUDA(cudaGetLastError());
    }
    return SB_OK;
  }
  if (with_dot) {
    cg_spmv_dot_kernel<<<reducing_grid(n, kWarpsPerCta), kThreads, 0, s>>>(op.A.row_ptr, op.A.col_ind, op.values, d, z, n,
                                                                           S->scalars, S->partials, S->ticket);
  } else {
    csr_spmv_kernel<<<reducing_grid(n, kWarpsPerCta), kThreads, 0, s>>>(op.A.row_ptr, op.A.col_ind, op.values, d, z, n);
  }
  sb_count_launch();
  SB_CHECK_CUDA(cudaGetLastError());
  return SB_OK;
}

int poll(sb_cg_t S, cudaStream_t s)
{
  SB_CHECK_CUDA(cudaMemcpyAsync(S->host, S->scalars, sizeof(CgScalars), cudaMemcpyDeviceToHost, s));
  SB_CHECK_CUDA(cudaStreamSynchronize(s));
  return SB_OK;
}

int cg_solve(sb_cg_t S, const CgOperator& op, const double* dinv, const double* b, double* x, const sb_cg_options* opt,
             sb_cg_result* result, cudaStream_t s)
{
  SB_REQUIRE(S && opt && result, "sb_cg_solve: null solver / options / result");
  SB_REQUIRE(opt->max_iter >= 0 && opt->check_every >= 1, "sb_cg_solve: max_iter >= 0 and check_every >= 1 required");
  const int64_t n = S->n;
  *result         = sb_cg_result{};
  if (n == 0) {  // (B r, r) = 0 <= r0: converged at once (:769-775)
    result->converged = 1;
    return SB_OK;
  }
  SB_REQUIRE(b && x, "sb_cg_solve: null vector");
  const int vgrid = reducing_grid(n, kThreads);

  cg_reset_kernel<<<1, 1, 0, s>>>(S->scalars, opt->rel_tol, opt->abs_tol, opt->max_iter, S->ticket);
  sb_count_launch();
  SB_CHECK_CUDA(cudaGetLastError());
  if (opt->iterative_mode)
    if (int e = apply_operator(S, op, x, S->r, false, s)) return e;  // r = A x  (:727)
  cg_initial_residual_kernel<<<vgrid, kThreads, 0, s>>>(b, S->r, x, opt->iterative_mode, n);
  sb_count_launch();
  SB_CHECK_CUDA(cudaGetLastError());
  cg_start_kernel<<<vgrid, kThreads, 0, s>>>(S->r, dinv, S->z, S->d, n, S->scalars, S->partials, S->ticket);
  sb_count_launch();
  SB_CHECK_CUDA(cudaGetLastError());
  // the operator's own kernels cannot see the `done` flag, so the early exits of the start are read back before them
  if (int e = poll(S, s)) return e;
  if (!S->host->done) {
    if (int e = apply_operator(S, op, S->d, S->z, true, s)) return e;  // z = A d, den  (:777-794)
    // mfem enters the loop with i = 1 and leaves it through one of the exits inside; at most max(max_iter, 1) updates
    const int     updates = std::max(opt->max_iter, 1);
    const double* zr      = dinv ? S->z : S->r;
    for (int i = 1; i <= updates; i++) {
      cg_update_kernel<<<vgrid, kThreads, 0, s>>>(x, S->r, S->z, S->d, dinv, n, S->scalars, S->partials, S->ticket);
      sb_count_launch();
      SB_CHECK_CUDA(cudaGetLastError());
      if (i % opt->check_every == 0 || i == updates) {
        if (int e = poll(S, s)) return e;
        if (S->host->done) break;
      }
      cg_direction_kernel<<<vgrid, kThreads, 0, s>>>(S->d, zr, n, S->scalars);
      sb_count_launch();
      SB_CHECK_CUDA(cudaGetLastError());
      if (int e = apply_operator(S, op, S->d, S->z, true, s)) return e;
    }
    if (!S->host->done) {  // den == 0 after the last product, or the flag was set by a kernel after the last poll
      if (int e = poll(S, s)) return e;
    }
  }
  const CgScalars& h   = *S->host;
  result->converged    = h.converged;
  result->final_iter   = h.final_iter;
  result->initial_norm = h.initial_norm;
  result->final_norm   = h.final_norm;
  if (!h.done) {
    sb::set_error("sb_cg_solve: the iteration ended without a verdict (internal error)");
    return SB_ERR_INVALID;
  }
  return SB_OK;
}

}  // namespace

extern "C" int sb_cg_solve_csr(sb_cg_t S, sb_csr_t A, const double* values_dev, const double* diag_inv_dev,
                               const double* b_dev, double* x_dev, const sb_cg_options* opt, sb_cg_result* result,
                               void* stream)
{
  SB_REQUIRE(S != nullptr, "sb_cg_solve_csr: null solver");
  CgOperator op;
  if (int e = view_of(A, &op.A)) return e;
  SB_REQUIRE(op.A.num_rows == S->n, "sb_cg_solve_csr: the solver was created for a different size");
  SB_REQUIRE(values_dev || op.A.nnz == 0, "sb_cg_solve_csr: null values");
  op.values = values_dev;
  return cg_solve(S, op, diag_inv_dev, b_dev, x_dev, opt, result, as_stream(stream));
}

extern "C" int sb_cg_solve_operator(sb_cg_t S, sb_operator_apply apply, void* ctx, const double* diag_inv_dev,
                                    const double* b_dev, double* x_dev, const sb_cg_options* opt, sb_cg_result* result,
                                    void* stream)
{
  SB_REQUIRE(S != nullptr && apply != nullptr, "sb_cg_solve_operator: null solver / operator");
  CgOperator op;
  op.apply = apply;
  op.ctx   = ctx;
  return cg_solve(S, op, diag_inv_dev, b_dev, x_dev, opt, result, as_stream(stream));
}

// ---- GMRES(m) ---------------------------------------------------------------------------------------------------------
// mfem::GMRESSolver::Mult (mfem/linalg/solvers.cpp:981-1165; line numbers below refer to that file): restarted GMRES with
// modified Gram-Schmidt.  The Krylov basis, the work vectors and the Hessenberg column being built stay on the device: the
// chain  H(k,i) = (w, v_k);  w -= H(k,i) v_k  runs as one kernel per k that applies the previous projection and reduces
// the next inner product (finished by the last block to arrive), with no host round trip inside a column; the host reads
// the finished column once per iteration for the (m+1) x m Givens / back-substitution arithmetic, as the reference does.
namespace {

/// w -= hcol[k-1] v_prev (when k > 0), then hcol[k] = (w, v_k)
__global__ void __launch_bounds__(kThreads) gmres_project_kernel(double* __restrict__ w, const double* __restrict__ v_prev,
                                                                 const double* __restrict__ v_k, double* __restrict__ hcol,
                                                                 int k, int64_t n, double* __restrict__ partials,
                                                                 unsigned int* __restrict__ ticket)
{
  __shared__ double smem[kWarpsPerCta];
  const double      h_prev = (k > 0) ? hcol[k - 1] : 0.0;
  double            acc    = 0.0;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    double wi = w[i];
    if (k > 0) {
      wi   = wi + (-h_prev) * v_prev[i];  // w.Add(-H(k,i), *v[k])  (:1074)
      w[i] = wi;
    }
    acc += wi * v_k[i];  // Dot(w, *v[k])  (:1073)
  }
  const double part  = block_sum(acc, smem);
  double       total = 0.0;
  if (grid_sum_last_block(part, partials, ticket, smem, &total) && threadIdx.x == 0) {
    hcol[k] = total;
    __threadfence();
  }
}

/// w -= hcol[i] v_i, then hcol[i+1] = ||w||  (:1074, :1077)
__global__ void __launch_bounds__(kThreads) gmres_norm_kernel(double* __restrict__ w, const double* __restrict__ v_i,
                                                              double* __restrict__ hcol, int i, int64_t n,
                                                              double* __restrict__ partials,
                                                              unsigned int* __restrict__ ticket)
{
  __shared__ double smem[kWarpsPerCta];
  const double      h   = hcol[i];
  double            acc = 0.0;
  for (int64_t j = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; j < n; j += int64_t(gridDim.x) * blockDim.x) {
    const double wj = w[j] + (-h) * v_i[j];
    w[j]            = wj;
    acc += wj * wj;
  }
  const double part  = block_sum(acc, smem);
  double       total = 0.0;
  if (grid_sum_last_block(part, partials, ticket, smem, &total) && threadIdx.x == 0) {
    hcol[i + 1] = sqrt(total);
    __threadfence();
  }
}

/// v_next = w / hcol[i+1]  (:1080)
__global__ void gmres_normalize_kernel(const double* __restrict__ w, double* __restrict__ v_next,
                                       const double* __restrict__ hcol, int i, int64_t n)
{
  const double s = 1.0 / hcol[i + 1];
  for (int64_t j = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; j < n; j += int64_t(gridDim.x) * blockDim.x)
    v_next[j] = s * w[j];
}

/// r = M (b - r) or r = b - r (or M b / b when `have_Ax` is 0), out[0] = ||r||  (:1004-1027, :1119-1129)
__global__ void __launch_bounds__(kThreads) gmres_residual_kernel(const double* __restrict__ b, double* __restrict__ r,
                                                                  const double* __restrict__ dinv, int have_Ax, int64_t n,
                                                                  double* __restrict__ out, double* __restrict__ partials,
                                                                  unsigned int* __restrict__ ticket)
{
  __shared__ double smem[kWarpsPerCta];
  double            acc = 0.0;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    double ri = have_Ax ? b[i] - r[i] : b[i];
    if (dinv) ri = dinv[i] * ri;
    r[i] = ri;
    acc += ri * ri;
  }
  const double part  = block_sum(acc, smem);
  double       total = 0.0;
  if (grid_sum_last_block(part, partials, ticket, smem, &total) && threadIdx.x == 0) {
    *out = sqrt(total);
    __threadfence();
  }
}

/// v0 = r / beta  (:1056)
__global__ void gmres_first_vector_kernel(const double* __restrict__ r, double* __restrict__ v0, double inv_beta, int64_t n)
{
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    v0[i] = inv_beta * r[i];
}

/// w = D^-1 r
__global__ void scale_by_kernel(const double* __restrict__ dinv, const double* __restrict__ r, double* __restrict__ w,
                                int64_t n)
{
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    w[i] = dinv[i] * r[i];
}

/// x += sum_j y[j] v_j, j ascending  (Update, :960-980)
__global__ void gmres_update_kernel(double* __restrict__ x, const double* __restrict__ V, const double* __restrict__ y,
                                    int count, int64_t n)
{
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    double xi = x[i];
    for (int j = 0; j < count; j++) xi = xi + y[j] * V[int64_t(j) * n + i];
    x[i] = xi;
  }
}

__global__ void zero_kernel(double* __restrict__ x, int64_t n)
{
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) x[i] = 0.0;
}

}  // namespace

struct sb_gmres_s {
  int64_t       n = 0;
  int           m = 0;
  double*       V        = nullptr;  // [m + 1][n]
  double *      r = nullptr, *w = nullptr;
  double*       hcol     = nullptr;  // [m + 2] the column being built; [m + 2 ...] y of the update (m entries); then beta
  double*       partials = nullptr;
  unsigned int* ticket   = nullptr;
  double*       host     = nullptr;  // pinned, m + 2 doubles
};

extern "C" int sb_gmres_destroy(sb_gmres_t G)
{
  if (!G) return SB_OK;
  cudaFree(G->V);
  cudaFree(G->r);
  cudaFree(G->w);
  cudaFree(G->hcol);
  cudaFree(G->partials);
  cudaFree(G->ticket);
  if (G->host) cudaFreeHost(G->host);
  delete G;
  return SB_OK;
}

extern "C" int sb_gmres_create(int64_t n, int kdim, sb_gmres_t* out)
{
  SB_REQUIRE(out && n >= 0 && kdim >= 1, "sb_gmres_create: bad argument");
  auto G = std::make_unique<sb_gmres_s>();
  G->n   = n;
  G->m   = kdim;
  int rc = dev_new(&G->V, size_t(kdim + 1) * size_t(n));
  if (!rc) rc = dev_new(&G->r, size_t(n));
  if (!rc) rc = dev_new(&G->w, size_t(n));
  if (!rc) rc = dev_new(&G->hcol, size_t(2 * kdim + 4));
  if (!rc) rc = dev_new(&G->partials, size_t(kMaxPartials));
  if (!rc) rc = dev_new(&G->ticket, 1);
  if (!rc && cudaMemset(G->ticket, 0, sizeof(unsigned int)) != cudaSuccess) rc = SB_ERR_CUDA;
  if (!rc && cudaMallocHost(reinterpret_cast<void**>(&G->host), size_t(kdim + 2) * sizeof(double)) != cudaSuccess) {
    sb::set_error("sb_gmres_create: cudaMallocHost failed");
    rc = SB_ERR_CUDA;
  }
  if (rc) {
    sb_gmres_destroy(G.release());
    return rc;
  }
  *out = G.release();
  return SB_OK;
}

namespace {

int apply_plain(const CgOperator& op, int64_t n, const double* x, double* y, cudaStream_t s)
{
  if (op.apply) return op.apply(op.ctx, x, y, s);
  csr_spmv_kernel<<<reducing_grid(n, kWarpsPerCta), kThreads, 0, s>>>(op.A.row_ptr, op.A.col_ind, op.values, x, y, n);
  sb_count_launch();
  SB_CHECK_CUDA(cudaGetLastError());
  return SB_OK;
}

// GeneratePlaneRotation / ApplyPlaneRotation (:931-958)
void generate_rotation(double dx, double dy, double& cs, double& sn)
{
  if (dy == 0.0) {
    cs = 1.0;
    sn = 0.0;
  } else if (std::fabs(dy) > std::fabs(dx)) {
    const double t = dx / dy;
    sn             = 1.0 / std::sqrt(1.0 + t * t);
    cs             = t * sn;
  } else {
    const double t = dy / dx;
    cs             = 1.0 / std::sqrt(1.0 + t * t);
    sn             = t * cs;
  }
}
void apply_rotation(double& dx, double& dy, double cs, double sn)
{
  const double t = cs * dx + sn * dy;
  dy             = -sn * dx + cs * dy;
  dx             = t;
}

int gmres_solve(sb_gmres_t G, const CgOperator& op, const double* dinv, const double* b, double* x,
                const sb_cg_options* opt, sb_cg_result* result, cudaStream_t s)
{
  SB_REQUIRE(G && opt && result, "sb_gmres_solve: null solver / options / result");
  SB_REQUIRE(opt->max_iter >= 0, "sb_gmres_solve: max_iter >= 0 required");
  const int64_t n = G->n;
  const int     m = G->m;
  *result         = sb_cg_result{};
  if (n == 0) {
    result->converged = 1;
    return SB_OK;
  }
  SB_REQUIRE(b && x, "sb_gmres_solve: null vector");
  const int vgrid    = reducing_grid(n, kThreads);
  double*   beta_dev = G->hcol + 2 * m + 3;
  double*   y_dev    = G->hcol + m + 2;
  auto      v        = [&](int i) { return G->V + int64_t(i) * n; };

  // r = M (b - A x), beta = ||r||  (:995-1027)
  auto residual = [&](bool have_Ax, double* beta) -> int {
    gmres_residual_kernel<<<vgrid, kThreads, 0, s>>>(b, G->r, dinv, have_Ax ? 1 : 0, n, beta_dev, G->partials, G->ticket);
    sb_count_launch();
    SB_CHECK_CUDA(cudaGetLastError());
    SB_CHECK_CUDA(cudaMemcpyAsync(G->host, beta_dev, sizeof(double), cudaMemcpyDeviceToHost, s));
    SB_CHECK_CUDA(cudaStreamSynchronize(s));
    *beta = G->host[0];
    return SB_OK;
  };
  // x += sum_i y_i v_i with H y = s back-substituted on the host  (Update, :960-980)
  std::vector<double> H(size_t(m + 1) * size_t(m), 0.0), sv(size_t(m + 1)), cs(size_t(m + 1)), sn(size_t(m + 1));
  auto                Hat    = [&](int r_, int c_) -> double& { return H[size_t(r_) * size_t(m) + size_t(c_)]; };
  auto                update = [&](int k) -> int {
    if (k < 0) return SB_OK;
    std::vector<double> y(sv.begin(), sv.begin() + k + 1);
    for (int i = k; i >= 0; i--) {
      y[size_t(i)] /= Hat(i, i);
      for (int j = i - 1; j >= 0; j--) y[size_t(j)] -= Hat(j, i) * y[size_t(i)];
    }
    SB_CHECK_CUDA(cudaMemcpyAsync(y_dev, y.data(), size_t(k + 1) * sizeof(double), cudaMemcpyHostToDevice, s));
    gmres_update_kernel<<<vgrid, kThreads, 0, s>>>(x, G->V, y_dev, k + 1, n);
    sb_count_launch();
    SB_CHECK_CUDA(cudaGetLastError());
    SB_CHECK_CUDA(cudaStreamSynchronize(s));  // y is a stack vector of this call
    return SB_OK;
  };

  double beta = 0.0;
  if (opt->iterative_mode) {
    if (int e = apply_plain(op, n, x, G->r, s)) return e;
  } else {
    zero_kernel<<<vgrid, kThreads, 0, s>>>(x, n);
    sb_count_launch();
    SB_CHECK_CUDA(cudaGetLastError());
  }
  if (int e = residual(opt->iterative_mode != 0, &beta)) return e;
  result->initial_norm = beta;
  SB_REQUIRE(std::isfinite(beta), "sb_gmres_solve: the initial residual is not finite");
  double final_norm = std::max(opt->rel_tol * beta, opt->abs_tol);  // :1030
  if (beta <= final_norm) {                                         // :1032-1039
    result->final_norm = beta;
    result->final_iter = 0;
    result->converged  = 1;
    return SB_OK;
  }
  int j = 1;
  while (j <= opt->max_iter) {  // :1053
    gmres_first_vector_kernel<<<vgrid, kThreads, 0, s>>>(G->r, v(0), 1.0 / beta, n);
    sb_count_launch();
    SB_CHECK_CUDA(cudaGetLastError());
    std::fill(sv.begin(), sv.end(), 0.0);
    sv[0] = beta;
    int i = 0;
    for (; i < m && j <= opt->max_iter; i++, j++) {  // :1059
      if (dinv) {                                     // w = M A v_i  (:1061-1069)
        if (int e = apply_plain(op, n, v(i), G->r, s)) return e;
        scale_by_kernel<<<vgrid, kThreads, 0, s>>>(dinv, G->r, G->w, n);
        sb_count_launch();
        SB_CHECK_CUDA(cudaGetLastError());
      } else {
        if (int e = apply_plain(op, n, v(i), G->w, s)) return e;
      }
      for (int k = 0; k <= i; k++) {  // modified Gram-Schmidt  (:1071-1075)
        gmres_project_kernel<<<vgrid, kThreads, 0, s>>>(G->w, k > 0 ? v(k - 1) : nullptr, v(k), G->hcol, k, n, G->partials,
                                                        G->ticket);
        sb_count_launch();
      }
      gmres_norm_kernel<<<vgrid, kThreads, 0, s>>>(G->w, v(i), G->hcol, i, n, G->partials, G->ticket);
      sb_count_launch();
      gmres_normalize_kernel<<<vgrid, kThreads, 0, s>>>(G->w, v(i + 1), G->hcol, i, n);
      sb_count_launch();
      SB_CHECK_CUDA(cudaGetLastError());
      SB_CHECK_CUDA(cudaMemcpyAsync(G->host, G->hcol, size_t(i + 2) * sizeof(double), cudaMemcpyDeviceToHost, s));
      SB_CHECK_CUDA(cudaStreamSynchronize(s));
      for (int k = 0; k <= i + 1; k++) Hat(k, i) = G->host[k];
      SB_REQUIRE(std::isfinite(Hat(i + 1, i)), "sb_gmres_solve: Norm(w) is not finite");
      for (int k = 0; k < i; k++) apply_rotation(Hat(k, i), Hat(k + 1, i), cs[size_t(k)], sn[size_t(k)]);  // :1082-1085
      generate_rotation(Hat(i, i), Hat(i + 1, i), cs[size_t(i)], sn[size_t(i)]);
      apply_rotation(Hat(i, i), Hat(i + 1, i), cs[size_t(i)], sn[size_t(i)]);
      apply_rotation(sv[size_t(i)], sv[size_t(i) + 1], cs[size_t(i)], sn[size_t(i)]);
      const double resid = std::fabs(sv[size_t(i) + 1]);  // :1091
      SB_REQUIRE(std::isfinite(resid), "sb_gmres_solve: the residual estimate is not finite");
      if (resid <= final_norm) {  // :1094-1101
        if (int e = update(i)) return e;
        result->final_norm = resid;
        result->final_iter = j;
        result->converged  = 1;
        return SB_OK;
      }
    }
    if (int e = update(i - 1)) return e;                     // :1117
    if (int e = apply_plain(op, n, x, G->r, s)) return e;    // :1119
    if (int e = residual(true, &beta)) return e;             // :1120-1129
    SB_REQUIRE(std::isfinite(beta), "sb_gmres_solve: the residual is not finite");
    if (beta <= final_norm) {  // :1131-1137
      result->final_norm = beta;
      result->final_iter = j;
      result->converged  = 1;
      return SB_OK;
    }
  }
  result->final_norm = beta;  // :1140-1142
  result->final_iter = opt->max_iter;
  result->converged  = 0;
  return SB_OK;
}

}  // namespace

extern "C" int sb_gmres_solve_csr(sb_gmres_t G, sb_csr_t A, const double* values_dev, const double* diag_inv_dev,
                                  const double* b_dev, double* x_dev, const sb_cg_options* opt, sb_cg_result* result,
                                  void* stream)
{
  SB_REQUIRE(G != nullptr, "sb_gmres_solve_csr: null solver");
  CgOperator op;
  if (int e = view_of(A, &op.A)) return e;
  SB_REQUIRE(op.A.num_rows == G->n, "sb_gmres_solve_csr: the solver was created for a different size");
  SB_REQUIRE(values_dev || op.A.nnz == 0, "sb_gmres_solve_csr: null values");
  op.values = values_dev;
  return gmres_solve(G, op, diag_inv_dev, b_dev, x_dev, opt, result, as_stream(stream));
}

extern "C" int sb_gmres_solve_operator(sb_gmres_t G, sb_operator_apply apply, void* ctx, const double* diag_inv_dev,
                                       const double* b_dev, double* x_dev, const sb_cg_options* opt, sb_cg_result* result,
                                       void* stream)
{
  SB_REQUIRE(G != nullptr && apply != nullptr, "sb_gmres_solve_operator: null solver / operator");
  CgOperator op;
  op.apply = apply;
  op.ctx   = ctx;
  return gmres_solve(G, op, diag_inv_dev, b_dev, x_dev, opt, result, as_stream(stream));
}
