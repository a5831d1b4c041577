// serac_b200: the Jacobian-vector product as a persistent, TMA-fed streaming kernel.
//
// Same arithmetic as jvp_kernel (domain_integral.cuh; reference: action_of_gradient_kernel,
// src/serac/numerics/functional/domain_integral_kernels.hpp:277-304 with chain_rule :223-237), different data movement.
// With the reference's full derivative layout the kernel is HBM-bound on the stored derivatives (648 of 666 B per
// quadrature point for solids); with symmetric tangent storage (DerivativeLayout::sym, 360 B) it runs at the shared-memory
// pipe (profiles/r4_summary.md).  So:
//   * CTAs are persistent (grid = resident CTAs x SMs) and walk the derivative tiles round-robin;
//   * each tile -- [entry][lane] doubles, contiguous in HBM -- is brought into shared memory by ONE bulk async copy
//     (cp.async.bulk / TMA, completion on an mbarrier), issued a whole tile ahead of its use, so the copy engine keeps
//     tens of KB in flight per SM without costing registers;
//   * the du gather of tile i+1 (table -> L dependent loads) is issued into registers while tile i is contracted and
//     integrated, so its latency never sits on the critical path;
//   * the 1-D basis rows / columns every thread needs live in registers for the whole kernel (the sum-factorised
//     stages then cost one shared-memory load per FMA instead of two);
//   * a 27-point hexahedron is one warp's own (JvpStreamConfig::LPE): its scratch blocks start on 128-byte lines, the
//     stages exchange through __syncwarp, and only the tile buffer is shared by the CTA's warps.
#pragma once

#include <cstdint>

#include "serac_b200/domain_integral.cuh"

namespace serac {
namespace b200 {

namespace tma {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
/// 1-D bulk copy global -> shared (SASS: UBLKCP); dst/src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
}  // namespace tma

/// this thread's rows of B and G (interpolation) and columns (integration) in each direction
template <int dim, int n, int q>
struct ThreadBasis {
  double Br[dim][n], Gr[dim][n];  ///< B(q_d, 0..n-1): row of this thread's quadrature index in direction d
  double Bc[dim][q], Gc[dim][q];  ///< B(0..q-1, min(q_d, n-1)): column of this thread's node index in direction d
  __device__ __forceinline__ void load(const double* __restrict__ B, const double* __restrict__ G, const int (&qi)[3])
  {
    SB_UNROLL
    for (int d = 0; d < dim; d++) {
      SB_UNROLL
      for (int k = 0; k < n; k++) {
        Br[d][k] = B[qi[d] * n + k];
        Gr[d][k] = G[qi[d] * n + k];
      }
      const int col = qi[d] < n ? qi[d] : n - 1;
      SB_UNROLL
      for (int k = 0; k < q; k++) {
        Bc[d][k] = B[k * n + col];
        Gc[d][k] = G[k * n + col];
      }
    }
  }
};

/// exchange point between two stages: the CTA's barrier, or the warp's when an element is one warp's own
template <bool WS>
__device__ __forceinline__ void jvp_stage_sync()
{
  if constexpr (WS)
    __syncwarp();
  else
    __syncthreads();
}

/// sum-factorised interpolation / integration with the 1-D tables in registers (same staging and summation order as
/// TensorProduct<> in domain_integral.cuh, hence the same bits)
template <int dim, int n, int q, int c>
struct TensorProductR;

template <int n, int q, int c>
struct TensorProductR<3, n, q, c> {
  using TP = TensorProduct<3, n, q, c>;
  // Every intermediate array is a set of blocks of (at most) q^3 doubles, one per (kind, component), each block starting
  // on a 128-byte line (BLK doubles apart): a warp's access to one block then touches the fewest lines there are.  Inside
  // a block the index that a stage sums over is the slowest, so that the lanes of one load read one contiguous run
  // (ncu on the [qz][dy][qx] order: 3 wavefronts per load in the y stages instead of 2; profiles/r4_summary.md).
  static constexpr int BLK            = ((q * q * q + 15) / 16) * 16;
  static constexpr int U_SZ           = ((c * n * n * n + 15) / 16) * 16;
  static constexpr int interp_scratch = U_SZ + (2 + 3) * c * BLK;
  static constexpr int integ_scratch  = (4 + 3 + 2) * c * BLK;
  /// WS: the element is one warp's own (JvpStreamConfig::LPE == 32): the stages exchange through __syncwarp
  template <bool NEED_VALUE, bool NEED_GRAD, bool WS = false>
  static __device__ __forceinline__ void interpolate(const ThreadBasis<3, n, q>& tb, double* s, int qx, int qy, int qz,
                                                     double (&val)[c], double (&grad)[c][3], bool on = true)
  {
    const double* u    = s;
    double*       A1   = s + U_SZ;           // blocks (kind < 2, a): [dy][qz][qx]
    double*       A2   = A1 + 2 * c * BLK;   // blocks (kind < 3, a): [dz][qy][qx]
    if (on && qz < n && qy < n) {
      SB_UNROLL
      for (int a = 0; a < c; a++) {
        double sB = 0.0, sG = 0.0;
        SB_UNROLL
        for (int dx = 0; dx < n; dx++) {
          const double uu = u[((a * n + qz) * n + qy) * n + dx];
          sB += tb.Br[0][dx] * uu;
          if (NEED_GRAD) sG += tb.Gr[0][dx] * uu;
        }
        A1[(0 * c + a) * BLK + (qy * n + qz) * q + qx] = sB;
        if (NEED_GRAD) A1[(1 * c + a) * BLK + (qy * n + qz) * q + qx] = sG;
      }
    }
    jvp_stage_sync<WS>();
    if (on && qz < n) {
      SB_UNROLL
      for (int a = 0; a < c; a++) {
        double sBB = 0.0, sGB = 0.0, sBG = 0.0;
        SB_UNROLL
        for (int dy = 0; dy < n; dy++) {
          const double b = A1[(0 * c + a) * BLK + (dy * n + qz) * q + qx];
          sBB += tb.Br[1][dy] * b;
          if (NEED_GRAD) {
            const double g = A1[(1 * c + a) * BLK + (dy * n + qz) * q + qx];
            sGB += tb.Br[1][dy] * g;
            sBG += tb.Gr[1][dy] * b;
          }
        }
        A2[(0 * c + a) * BLK + (qz * q + qy) * q + qx] = sBB;
        if (NEED_GRAD) {
          A2[(1 * c + a) * BLK + (qz * q + qy) * q + qx] = sGB;
          A2[(2 * c + a) * BLK + (qz * q + qy) * q + qx] = sBG;
        }
      }
    }
    jvp_stage_sync<WS>();
    SB_UNROLL
    for (int a = 0; a < c; a++) {
      double v = 0.0, gx = 0.0, gy = 0.0, gz = 0.0;
      SB_UNROLL
      for (int dz = 0; dz < n; dz++) {
        const double bb = A2[(0 * c + a) * BLK + (dz * q + qy) * q + qx];
        if (NEED_VALUE) v += tb.Br[2][dz] * bb;
        if (NEED_GRAD) {
          const double gb = A2[(1 * c + a) * BLK + (dz * q + qy) * q + qx];
          const double bg = A2[(2 * c + a) * BLK + (dz * q + qy) * q + qx];
          gx += tb.Br[2][dz] * gb;
          gy += tb.Br[2][dz] * bg;
          gz += tb.Gr[2][dz] * bb;
        }
      }
      val[a]     = v;
      grad[a][0] = gx;
      grad[a][1] = gy;
      grad[a][2] = gz;
    }
  }

  template <bool HAS_SOURCE, bool HAS_FLUX, bool WS = false>
  static __device__ __forceinline__ bool integrate(const ThreadBasis<3, n, q>& tb, double* s, int qx, int qy, int qz,
                                                   const double (&src)[c], const double (&flux)[c][3],
                                                   double (&out)[c], bool on = true)
  {
    double* S  = s;                  // blocks (kind < 4, a): [qz][qy][qx]
    double* P  = S + 4 * c * BLK;    // blocks (kind < 3, a): [qy][jz][qx]
    double* Qs = P + 3 * c * BLK;    // blocks (kind < 2, a): [jz][jy][qx]
    SB_UNROLL
    for (int a = 0; a < c; a++) {
      if (HAS_SOURCE && on) S[(0 * c + a) * BLK + (qz * q + qy) * q + qx] = src[a];
      if (HAS_FLUX && on) {
        S[(1 * c + a) * BLK + (qz * q + qy) * q + qx] = flux[a][0];
        S[(2 * c + a) * BLK + (qz * q + qy) * q + qx] = flux[a][1];
        S[(3 * c + a) * BLK + (qz * q + qy) * q + qx] = flux[a][2];
      }
    }
    jvp_stage_sync<WS>();
    if (on && qz < n) {
      SB_UNROLL
      for (int a = 0; a < c; a++) {
        double p0 = 0.0, p1 = 0.0, p2 = 0.0;
        SB_UNROLL
        for (int k = 0; k < q; k++) {
          if (HAS_SOURCE) p0 += tb.Bc[2][k] * S[(0 * c + a) * BLK + (k * q + qy) * q + qx];
          if (HAS_FLUX) {
            p0 += tb.Gc[2][k] * S[(3 * c + a) * BLK + (k * q + qy) * q + qx];
            p1 += tb.Bc[2][k] * S[(1 * c + a) * BLK + (k * q + qy) * q + qx];
            p2 += tb.Bc[2][k] * S[(2 * c + a) * BLK + (k * q + qy) * q + qx];
          }
        }
        P[(0 * c + a) * BLK + (qy * n + qz) * q + qx] = p0;
        if (HAS_FLUX) {
          P[(1 * c + a) * BLK + (qy * n + qz) * q + qx] = p1;
          P[(2 * c + a) * BLK + (qy * n + qz) * q + qx] = p2;
        }
      }
    }
    jvp_stage_sync<WS>();
    if (on && qz < n && qy < n) {
      SB_UNROLL
      for (int a = 0; a < c; a++) {
        double q0 = 0.0, q1 = 0.0;
        SB_UNROLL
        for (int k = 0; k < q; k++) {
          q0 += tb.Bc[1][k] * P[(0 * c + a) * BLK + (k * n + qz) * q + qx];
          if (HAS_FLUX) {
            q0 += tb.Gc[1][k] * P[(2 * c + a) * BLK + (k * n + qz) * q + qx];
            q1 += tb.Bc[1][k] * P[(1 * c + a) * BLK + (k * n + qz) * q + qx];
          }
        }
        Qs[(0 * c + a) * BLK + (qz * n + qy) * q + qx] = q0;
        if (HAS_FLUX) Qs[(1 * c + a) * BLK + (qz * n + qy) * q + qx] = q1;
      }
    }
    jvp_stage_sync<WS>();
    const bool owner = (on && qz < n && qy < n && qx < n);
    if (owner) {
      SB_UNROLL
      for (int a = 0; a < c; a++) {
        double r = 0.0;
        SB_UNROLL
        for (int k = 0; k < q; k++) {
          r += tb.Bc[0][k] * Qs[(0 * c + a) * BLK + (qz * n + qy) * q + k];
          if (HAS_FLUX) r += tb.Gc[0][k] * Qs[(1 * c + a) * BLK + (qz * n + qy) * q + k];
        }
        out[a] = r;
      }
    }
    return owner;
  }
};

template <int n, int q, int c>
struct TensorProductR<2, n, q, c> {
  template <bool NEED_VALUE, bool NEED_GRAD, bool WS = false>
  static __device__ __forceinline__ void interpolate(const ThreadBasis<2, n, q>& tb, double* s, int qx, int qy, int,
                                                     double (&val)[c], double (&grad)[c][2], bool = true)
  {
    constexpr int ndof = n * n;
    const double* u    = s;
    double*       A1   = s + c * ndof;
    if (qy < n) {
      SB_UNROLL
      for (int a = 0; a < c; a++) {
        double sB = 0.0, sG = 0.0;
        SB_UNROLL
        for (int dx = 0; dx < n; dx++) {
          const double uu = u[(a * n + qy) * n + dx];
          sB += tb.Br[0][dx] * uu;
          if (NEED_GRAD) sG += tb.Gr[0][dx] * uu;
        }
        A1[((0 * c + a) * n + qy) * q + qx] = sB;
        if (NEED_GRAD) A1[((1 * c + a) * n + qy) * q + qx] = sG;
      }
    }
    __syncthreads();
    SB_UNROLL
    for (int a = 0; a < c; a++) {
      double v = 0.0, gx = 0.0, gy = 0.0;
      SB_UNROLL
      for (int dy = 0; dy < n; dy++) {
        const double b = A1[((0 * c + a) * n + dy) * q + qx];
        if (NEED_VALUE) v += tb.Br[1][dy] * b;
        if (NEED_GRAD) {
          gx += tb.Br[1][dy] * A1[((1 * c + a) * n + dy) * q + qx];
          gy += tb.Gr[1][dy] * b;
        }
      }
      val[a]     = v;
      grad[a][0] = gx;
      grad[a][1] = gy;
    }
  }

  template <bool HAS_SOURCE, bool HAS_FLUX, bool WS = false>
  static __device__ __forceinline__ bool integrate(const ThreadBasis<2, n, q>& tb, double* s, int qx, int qy, int,
                                                   const double (&src)[c], const double (&flux)[c][2],
                                                   double (&out)[c], bool = true)
  {
    double* S = s;
    double* P = S + 3 * c * q * q;
    SB_UNROLL
    for (int a = 0; a < c; a++) {
      if (HAS_SOURCE) S[((0 * c + a) * q + qy) * q + qx] = src[a];
      if (HAS_FLUX) {
        S[((1 * c + a) * q + qy) * q + qx] = flux[a][0];
        S[((2 * c + a) * q + qy) * q + qx] = flux[a][1];
      }
    }
    __syncthreads();
    if (qy < n) {
      SB_UNROLL
      for (int a = 0; a < c; a++) {
        double p0 = 0.0, p1 = 0.0;
        SB_UNROLL
        for (int k = 0; k < q; k++) {
          if (HAS_SOURCE) p0 += tb.Bc[1][k] * S[((0 * c + a) * q + k) * q + qx];
          if (HAS_FLUX) {
            p0 += tb.Gc[1][k] * S[((2 * c + a) * q + k) * q + qx];
            p1 += tb.Bc[1][k] * S[((1 * c + a) * q + k) * q + qx];
          }
        }
        P[((0 * c + a) * n + qy) * q + qx] = p0;
        if (HAS_FLUX) P[((1 * c + a) * n + qy) * q + qx] = p1;
      }
    }
    __syncthreads();
    const bool owner = (qy < n && qx < n);
    if (owner) {
      SB_UNROLL
      for (int a = 0; a < c; a++) {
        double r = 0.0;
        SB_UNROLL
        for (int k = 0; k < q; k++) {
          r += tb.Bc[0][k] * P[((0 * c + a) * n + qy) * q + k];
          if (HAS_FLUX) r += tb.Gc[0][k] * P[((1 * c + a) * n + qy) * q + k];
        }
        out[a] = r;
      }
    }
    return owner;
  }
};

template <int dim, int q, int nk, int ck, int nt, int ct, int MASK, int EPB_ = 0>
struct JvpStreamConfig {
  using Base                     = JvpConfig<dim, q, nk, ck, nt, ct, MASK, EPB_>;
  static constexpr int QPE       = Base::QPE;
  static constexpr int EPB       = Base::EPB;
  /// lanes per element: a hexahedron with 27 points gets a warp of its own (5 idle lanes).  With 27 lanes per element
  /// the elements of a CTA straddle its warps, and every shared-memory access of the straddling warp touches two
  /// elements' blocks: 2.5 wavefronts on average instead of 2 (ncu source page, profiles/r4_summary.md)
  static constexpr int LPE       = (dim == 3 && Base::DQ_TE == 1 && QPE > 16 && QPE < 32) ? 32 : QPE;
  static constexpr int threads   = EPB * LPE;
  /// scratch doubles per element: the register-basis contraction of hexahedra pads its arrays to 128-byte lines
  static constexpr int stream_per()
  {
    if constexpr (dim == 3) {
      const int a = TensorProductR<3, nk, q, ck>::interp_scratch, b = TensorProductR<3, nt, q, ct>::integ_scratch;
      return a > b ? a : b;
    } else {
      return Base::per;
    }
  }
  static constexpr int per       = stream_per();
  static constexpr int entries   = Base::layout().entries;
  static constexpr int tile_doubles = entries * EPB * QPE;
  /// the bulk copy moves whole 16-byte units
  static constexpr bool supported = (tile_doubles % 2 == 0) && (tile_doubles * 8 < (1 << 20));
  static constexpr int  ndof_k    = ipow(nk, dim);
  static constexpr int  gathers   = EPB * ck * ndof_k;
  /// an element per warp: the warp gathers its own element's du and nothing but the tile buffer is shared in the CTA
  static constexpr bool warp_owned = (LPE == 32);
  static constexpr int  NG = warp_owned ? (ck * ndof_k + 31) / 32 : (gathers + threads - 1) / threads;  ///< du values each thread prefetches
  // smem: [tables 4 x MAX_Q^2][scratch EPB x per (+1 pad to keep the tile 16-byte aligned)][tile][mbarrier]
  static constexpr int scratch_doubles = ((4 * MAX_Q * MAX_Q + EPB * per + 1) / 2) * 2;
  static constexpr size_t smem_bytes   = size_t(scratch_doubles + tile_doubles) * 8 + 16;
};

#ifndef SB_JVP_MINB
#define SB_JVP_MINB 2
#endif
/// MINB: resident CTAs per SM asked of ptxas (part of the kernel's name so that tuning variants can live side by side)
template <int dim, int q, int nk, int ck, int nt, int ct, int MASK, int EPB_ = 0, int MINB = SB_JVP_MINB>
__global__ void __launch_bounds__(JvpStreamConfig<dim, q, nk, ck, nt, ct, MASK, EPB_>::threads, MINB)
    jvp_stream_kernel(const __grid_constant__ JvpParams P, const int num_tiles)
{
  using C                       = JvpStreamConfig<dim, q, nk, ck, nt, ct, MASK, EPB_>;
  using B                       = typename C::Base;
  constexpr DerivativeLayout LY = B::layout();
  constexpr int QPE = C::QPE, EPB = C::EPB, per = C::per, T = C::threads, NG = C::NG, ndk = C::ndof_k;
  extern __shared__ __align__(128) double smem[];
  double*   tables  = smem;
  double*   scratch = smem + 4 * MAX_Q * MAX_Q;
  double*   tile_s  = smem + C::scratch_doubles;
  uint64_t* bar     = reinterpret_cast<uint64_t*>(tile_s + C::tile_doubles);
  const int tid     = threadIdx.x;

  for (int i = tid; i < 4 * MAX_Q * MAX_Q; i += T) {
    const int           which = i / (2 * MAX_Q * MAX_Q);
    const int           rem   = i - which * (2 * MAX_Q * MAX_Q);
    const BasisTable1D& bt    = (which == 0) ? P.trial_basis : P.test_basis;
    tables[i]                 = (rem < MAX_Q * MAX_Q) ? bt.B[rem] : bt.G[rem - MAX_Q * MAX_Q];
  }
  if (tid == 0) tma::mbar_init(bar, 1);
  __syncthreads();

  constexpr uint32_t tile_bytes = uint32_t(C::tile_doubles) * 8u;
  int                tile       = blockIdx.x;
  if (tid == 0 && tile < num_tiles) {
    tma::mbar_expect_tx(bar, tile_bytes);
    tma::bulk_g2s(tile_s, P.dqf + (long long)tile * C::tile_doubles, tile_bytes, bar);
  }

  const int  el  = tid / C::LPE;
  const int  ql0 = tid - el * C::LPE;
  const bool on  = ql0 < QPE;               // (idle lanes shadow the element's last point and store nothing)
  const int  ql  = on ? ql0 : QPE - 1;
  const int qi[3] = {ql % q, (ql / q) % q, (dim == 3) ? ql / (q * q) : 0};
  const int qx = qi[0], qy = qi[1], qz = qi[2];

  ThreadBasis<dim, nk, q> tbk;
  ThreadBasis<dim, nt, q> tbt;
  tbk.load(tables, tables + MAX_Q * MAX_Q, qi);
  tbt.load(tables + 2 * MAX_Q * MAX_Q, tables + 3 * MAX_Q * MAX_Q, qi);
  const BasisTable1D& wt = P.test_basis;
  double              w  = wt.w[qx] * wt.w[qy];
  if (dim == 3) w *= wt.w[qz];

  // tile-invariant part of this thread's gather slots
  const FieldArg& f = P.trial;
  int             g_le[NG], g_node[NG];
  long long       g_comp[NG];
  int             g_slot[NG];  ///< position in the element's [component][node] block, -1 = none
  SB_UNROLL
  for (int j = 0; j < NG; j++) {
    int rem;
    if constexpr (C::warp_owned) {
      const int i = ql0 + j * 32;
      g_le[j]     = el;
      g_slot[j]   = (i < ck * ndk) ? i : -1;
      rem         = (i < ck * ndk) ? i : 0;
    } else {
      const int i  = tid + j * T;
      const int ii = (i < C::gathers) ? i : 0;
      g_le[j]      = ii / (ck * ndk);
      rem          = ii - g_le[j] * (ck * ndk);
      g_slot[j]    = (i < C::gathers) ? rem : -1;
    }
    const int a = rem / ndk;
    g_node[j]   = rem - a * ndk;
    g_comp[j]   = a * f.comp_stride;
  }
  auto gather = [&](int t, double(&v)[NG]) {
    SB_UNROLL
    for (int j = 0; j < NG; j++) {
      int e = t * EPB + g_le[j];
      if (e >= P.num_elements) e = P.num_elements - 1;
      const int       ge  = P.elements ? P.elements[e] : e;
      const long long idx = (long long)f.table[(long long)ge * ndk + g_node[j]];
      v[j]                = f.L[idx * f.node_stride + g_comp[j]];
    }
  };

  double nxt[NG];
  if (tile < num_tiles) gather(tile, nxt);
  uint32_t parity = 0;

  using TPk = TensorProductR<dim, nk, q, ck>;
  using TPt = TensorProductR<dim, nt, q, ct>;
  constexpr bool has_source = B::d00 || B::d01;
  constexpr bool has_flux   = B::d10 || B::d11;

  for (; tile < num_tiles; tile += gridDim.x) {
    constexpr bool WS = C::warp_owned;
    jvp_stage_sync<WS>();  // previous tile's integrate has finished reading the scratch
    SB_UNROLL
    for (int j = 0; j < NG; j++)
      if (g_slot[j] >= 0) scratch[g_le[j] * per + g_slot[j]] = nxt[j];
    jvp_stage_sync<WS>();
    double v[ck];
    double g[ck][dim];
    TPk::template interpolate<(B::d00 || B::d10), (B::d01 || B::d11), WS>(tbk, scratch + el * per, qx, qy, qz, v, g, on);

    const int next = tile + gridDim.x;
    if (next < num_tiles) gather(next, nxt);  // in flight while this tile is contracted and integrated

    tma::mbar_wait(bar, parity);
    parity ^= 1u;

    // lanes of one entry: the CTA's threads ([entry][threads] tiles) or this element's points ([element][entry][QPE])
    constexpr int DQS = B::DQ_STRIDE;
    const double* dq  = (B::DQ_TE == 1) ? tile_s + el * (LY.entries * C::QPE) + ql : tile_s + tid;
    double        src[ct];
    double        flux[ct][dim];
    SB_UNROLL
    for (int a = 0; a < ct; a++) {
      double s = 0.0;
      if constexpr (B::d00) {
        SB_UNROLL
        for (int b = 0; b < ck; b++) s += dq[(LY.offset[0] + a * ck + b) * DQS] * v[b];
      }
      if constexpr (B::d01) {
        SB_UNROLL
        for (int b = 0; b < ck; b++) {
          SB_UNROLL
          for (int l = 0; l < dim; l++) s += dq[(LY.offset[1] + (a * ck + b) * dim + l) * DQS] * g[b][l];
        }
      }
      src[a] = s * w;
      SB_UNROLL
      for (int k = 0; k < dim; k++) {
        double fl = 0.0;
        if constexpr (B::d10) {
          SB_UNROLL
          for (int b = 0; b < ck; b++) fl += dq[(LY.offset[2] + (a * dim + k) * ck + b) * DQS] * v[b];
        }
        if constexpr (B::d11) {
          SB_UNROLL
          for (int b = 0; b < ck; b++) {
            SB_UNROLL
            for (int l = 0; l < dim; l++) fl += dq[(LY.i11(a, k, b, l)) * DQS] * g[b][l];
          }
        }
        flux[a][k] = fl * w;
      }
    }
    __syncthreads();  // every thread is done with the tile buffer (and with the interpolation scratch)
    if (tid == 0 && next < num_tiles) {
      tma::mbar_expect_tx(bar, tile_bytes);
      tma::bulk_g2s(tile_s, P.dqf + (long long)next * C::tile_doubles, tile_bytes, bar);
    }
    double     out[ct];
    const bool owner = TPt::template integrate<has_source, has_flux, WS>(tbt, scratch + el * per, qx, qy, qz, src, flux, out, on);
    const int  e_raw = tile * EPB + el;
    if (owner && e_raw < P.num_elements) {
      const int elem = P.elements ? P.elements[e_raw] : e_raw;
      const int node = (dim == 3) ? (qz * nt + qy) * nt + qx : qy * nt + qx;
      SB_UNROLL
      for (int a = 0; a < ct; a++) P.r_E[e_index(P.node_major, elem, a, node, ct, ipow(nt, dim))] = out[a];
    }
  }
}

}  // namespace b200
}  // namespace serac
