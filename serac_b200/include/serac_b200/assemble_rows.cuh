// serac_b200: element-stiffness assembly fused into the CSR values -- no element matrices in HBM.
//
// Reference: Functional::Gradient::assemble (src/serac/numerics/functional/functional.hpp:586-644) computes every
// K_e (domain_integral_kernels.hpp:327-361) into an [ne][ndof][ndof] array and adds its entries into `values` through
// a hash lookup per entry.  Here the loop is turned around: the ROWS of the matrix are the unit of work.
//
//   * a group (host side, assembly.cu) = one test component `a` of the test nodes whose lowest-numbered element is the
//     same, cut so that the group's CSR rows fit in shared memory; ONE WARP (a CTA of 32 threads) owns a group and
//     therefore whole CSR rows -- nothing else touches them: no atomics, no read-modify-write of `values`, no block
//     barrier anywhere in the kernel;
//   * the unit of work is an item (element p, local test node I, block of trial components): the stored derivative is
//     contracted, at every quadrature point, with the test function  (phi_I, d phi_I/d xi)  and the result integrated
//     against ALL trial functions by sum factorisation (the transpose of what the reference does per trial node with
//     batch_apply_shape_fn + integrate, hexahedron_H1.inl:132-167,230-280); lane = (trial component, column of q
//     quadrature points); the two exchanges of the sum factorisation go through shared memory with __syncwarp;
//   * items are walked element-major, so a lane's slice of the element's derivatives is loaded once per element and
//     stays in registers for all test nodes of the group in that element; every row sees its elements in ascending
//     order -- the order of the reference's loop over elements;
//   * the rows accumulate in shared memory and are written to `values` once, as whole coalesced rows.
//
// The derivative entries of a lane's column live in registers for as long as the warp stays on one (element, a); trial
// basis tables are kernel parameters read with compile-time indices (uniform-register operands of DFMA).
#pragma once

#include <cstdint>

#include "serac_b200/domain_integral.cuh"

namespace serac {
namespace b200 {

constexpr int ASM_MAX_GROUP_NODES = 32;

struct AsmUnit {
  int p;       ///< position of the element in the integral's domain (indexes the stored derivatives)
  int meta;    ///< local test node I (bits 0-7) | trial block (8-15) | row slot in the group (16-23) | next block (24-31)
  int next_p;   ///< element of the next (element, block) the warp visits (-1: none): loaded into registers meanwhile
  int next2_p;  ///< the element after that (-1: none): prefetched into L2
};

struct AsmNode {
  int node;     ///< test node id
  int acc_off;  ///< offset of this node's row in the accumulator
  int rowlen;   ///< trial nodes in the row
  int row_start;  ///< row_ptr of the group's row of this node
};

struct AsmGroup {
  int node_begin, node_end;
  int unit_begin, unit_end;
  int acc_total;  ///< doubles of accumulator this group uses
  int a;          ///< test component of the group's rows
  int pad[2];
};

struct AssembleRowsParams {
  BasisTable1D          trial_basis;
  BasisTable1D          test_basis;
  const double*         dqf;     ///< stored derivatives, [tile][entry][lane]
  const AsmGroup*       groups;
  const AsmUnit*        units;
  const AsmNode*        nodes;
  const unsigned short* pos;     ///< [p][I][J]: position of trial node J in the row of test node I
  const int*            row_ptr; ///< vdof CSR
  double*               values;
  long long             row_ns, row_cs;  ///< test vdof = node * row_ns + comp * row_cs
  int                   col_byvdim;      ///< trial ordering: 1 = byVDIM (col = node * ck + b), 0 = byNODES
  int                   accumulate;      ///< 0: values = rows; 1: values += rows (second and later integrals)
  int                   num_groups;
};

template <int q, int nk, int ck, int nt, int ct, int MASK, int EPB>
struct AssembleRowsConfig {
  static constexpr int  dim = 3;
  static constexpr int  QPE = q * q * q, QQ = q * q;
  static constexpr int  TE   = dq_tile_elements(false, QPE, EPB);  ///< elements per stored-derivative tile
  static constexpr int  TILE = TE * QPE;
  static constexpr int  CPW  = (32 / QQ < ck) ? 32 / QQ : ck;  ///< trial components a warp handles at once
  static constexpr int  NB   = (ck + CPW - 1) / CPW;           ///< trial blocks
  static constexpr int  active  = CPW * QQ;                    ///< working lanes of a warp
  static constexpr int  threads = 32;
  /// resident warps per SM asked of ptxas: shared memory allows ~12 (p = 2 solids), which leaves 170 registers per lane --
  /// room for the next element's derivatives beside the current one's
  static constexpr int  min_blocks = 12;
  static constexpr int  UCAP = 96;  ///< items of a group staged in shared memory (16 bytes each)
  static constexpr bool d00 = (MASK & 1) != 0, d01 = (MASK & 2) != 0, d10 = (MASK & 4) != 0, d11 = (MASK & 8) != 0;
  static constexpr bool has_source = d00 || d10;  ///< multiplies phi_J
  static constexpr bool has_flux   = d01 || d11;  ///< multiplies d phi_J / d xi
  static constexpr int  NDT = nt * nt * nt, NDK = nk * nk * nk;
  /// scratch strides chosen so that a warp's 8-byte accesses fall into distinct banks in the producing AND the consuming
  /// stage: a stage-y reader (cl, jz, qx) at address jz * SP + cl * QQ + qy * q + qx needs SP = q (mod 16); a stage-x
  /// reader (cl, jz, jy) at jy * SQ + cl * QQ + jz * q + qx needs SQ = 1 (mod 16) when nk == q
  static constexpr int pad_to(int n, int r) { return n + ((r - n % 16) % 16 + 16) % 16; }
  static constexpr int  SP  = pad_to(active, q % 16);
  static constexpr int  SQ  = pad_to(active, 1);
  static constexpr int  PS  = 3 * nk * SP;   ///< stage-z results  [3][nk (jz)][SP: (cl, qy, qx)]
  static constexpr int  QS  = 2 * nk * SQ;   ///< stage-y results  [2][nk (jy)][SQ: (cl, jz, qx)]
  static constexpr int  fixed_doubles = 4 * MAX_Q * MAX_Q + 2 * ASM_MAX_GROUP_NODES + 2 * UCAP + PS + QS;
  static constexpr int  NE  = (d00 ? 1 : 0) + (d01 ? 3 : 0) + (d10 ? 3 : 0) + (d11 ? 9 : 0);  ///< entries per (a,b), point
  static constexpr DerivativeLayout layout() { return make_layout(ct, ck, dim, d00, d01, d10, d11, (MASK & SB_MASK_SYM) != 0); }
  static_assert(nk <= q && nt <= q, "quadrature rule must have at least p + 1 points per direction");
  static_assert(QQ <= 32, "a column set must fit a warp");
  static_assert(NB <= 255, "item encoding");
};

template <int q, int nk, int ck, int nt, int ct, int MASK, int EPB>
__global__ void __launch_bounds__(AssembleRowsConfig<q, nk, ck, nt, ct, MASK, EPB>::threads,
                                  AssembleRowsConfig<q, nk, ck, nt, ct, MASK, EPB>::min_blocks)
    assemble_rows_kernel(const __grid_constant__ AssembleRowsParams P)
{
  using C                       = AssembleRowsConfig<q, nk, ck, nt, ct, MASK, EPB>;
  constexpr DerivativeLayout LY = C::layout();
  constexpr int              QQ = C::QQ, SP = C::SP, SQ = C::SQ;

  extern __shared__ double smem[];
  double*  Bt    = smem;
  double*  Gt    = Bt + MAX_Q * MAX_Q;
  AsmNode* nodes = reinterpret_cast<AsmNode*>(Gt + 3 * MAX_Q * MAX_Q);  // ASM_MAX_GROUP_NODES x 16 bytes
  AsmUnit* items = reinterpret_cast<AsmUnit*>(Gt + 3 * MAX_Q * MAX_Q + 2 * ASM_MAX_GROUP_NODES);  // UCAP x 16 bytes
  double*  work  = Gt + 3 * MAX_Q * MAX_Q + 2 * ASM_MAX_GROUP_NODES + 2 * C::UCAP;
  double*  acc   = work + C::PS + C::QS;

  const AsmGroup grp = P.groups[blockIdx.x];

  double* Btw = Gt + MAX_Q * MAX_Q;  // B, G of the test space times the quadrature weight of the row's point
  double* Gtw = Btw + MAX_Q * MAX_Q;
  for (int i = threadIdx.x; i < MAX_Q * MAX_Q; i += C::threads) {
    Bt[i]  = P.test_basis.B[i];
    Gt[i]  = P.test_basis.G[i];
    Btw[i] = P.test_basis.B[i] * P.test_basis.w[(i / nt) < q ? (i / nt) : 0];
    Gtw[i] = P.test_basis.G[i] * P.test_basis.w[(i / nt) < q ? (i / nt) : 0];
  }
  for (int i = threadIdx.x; i < grp.node_end - grp.node_begin; i += C::threads) nodes[i] = P.nodes[grp.node_begin + i];
  for (int i = threadIdx.x; i < grp.unit_end - grp.unit_begin; i += C::threads)  // <= UCAP items per group (assembly.cu)
    items[i] = P.units[grp.unit_begin + i];
  for (int i = threadIdx.x; i < grp.acc_total; i += C::threads) acc[i] = 0.0;

  const int  lane   = threadIdx.x;
  const int  a      = grp.a;
  const int  cl     = lane / QQ;         // trial component within the block
  const int  col    = lane - cl * QQ;    // column index, re-interpreted per stage
  double*    Ps     = work;
  double*    Qs     = Ps + C::PS;
  // stage z: this lane's column of quadrature points (qy, qx)
  const int qy0 = col / q, qx0 = col - qy0 * q;
  // stage y: (jz, qx)
  const int jz1 = col / q, qx1 = col - jz1 * q;
  // stage x: (jz, jy)
  const int jz2 = col / nk, jy2 = col - jz2 * nk;

  // quadrature weights: w(qx) w(qy) goes into the test function's (x, y) factors, w(qz) into the staged z tables, so the
  // derivative entries are used as loaded
  const double wxy = P.test_basis.w[qx0 < q ? qx0 : 0] * P.test_basis.w[qy0 < q ? qy0 : 0];

  double D[q][C::NE > 0 ? C::NE : 1];   // derivative entries of (a,b) at the column's points
  double Dn[q][C::NE > 0 ? C::NE : 1];  // the same for the (element, block) the warp visits next, in flight
  long long cur_key = -1;               // (p, trial block) D holds
  __syncwarp();

  // this lane's slice of the stored derivative of element p, trial block bg
  auto load_D = [&](double (&dst)[q][C::NE > 0 ? C::NE : 1], int p, int bg) {
    const int b = bg * C::CPW + cl;
    if (cl < C::CPW && b < ck) {
      const long long tile = p / C::TE;
      const int       ln   = (p - int(tile) * C::TE) * C::QPE + qy0 * q + qx0;
      const double*   dq   = P.dqf + tile * (long long)LY.entries * C::TILE + ln;
      SB_UNROLL
      for (int qz = 0; qz < q; qz++) {
        int i = 0;
        if constexpr (C::d00) dst[qz][i++] = dq[(long long)(LY.offset[0] + a * ck + b) * C::TILE + qz * QQ];
        if constexpr (C::d01) {
          SB_UNROLL
          for (int l = 0; l < 3; l++) dst[qz][i++] = dq[(long long)(LY.offset[1] + (a * ck + b) * 3 + l) * C::TILE + qz * QQ];
        }
        if constexpr (C::d10) {
          SB_UNROLL
          for (int k = 0; k < 3; k++) dst[qz][i++] = dq[(long long)(LY.offset[2] + (a * 3 + k) * ck + b) * C::TILE + qz * QQ];
        }
        if constexpr (C::d11) {
          SB_UNROLL
          for (int k = 0; k < 3; k++) {
            SB_UNROLL
            for (int l = 0; l < 3; l++)
              dst[qz][i++] = dq[(long long)(LY.i11(a, k, b, l)) * C::TILE + qz * QQ];
          }
        }
      }
    }
  };
  // the rows of test component a of element p's derivative into L2: LY.entries / ct runs of QPE doubles
  auto prefetch_element = [&](int p) {
    const long long tile = p / C::TE;
    const double*   nq   = P.dqf + tile * (long long)LY.entries * C::TILE + (p - int(tile) * C::TE) * C::QPE;
    constexpr int   per_a = LY.entries / ct;  // entries are a-major inside every block; a few lines too many do no harm
    for (int e = lane; e < LY.entries; e += 32) {
      if (!LY.sym && ct > 1 && LY.entries % ct == 0 && !(C::d00 || C::d01 || C::d10) && e / per_a != a) continue;
      const char* at = reinterpret_cast<const char*>(nq + (long long)e * C::TILE);
      prefetch_l2(at);
      if (C::QPE * 8 > 128) prefetch_l2(at + 128);
      if (C::QPE * 8 > 256) prefetch_l2(at + 256);
      if (C::QPE * 8 > 384) prefetch_l2(at + 384);
      prefetch_l2(at + C::QPE * 8 - 8);
    }
  };

  const int n_items = grp.unit_end - grp.unit_begin;
  if (n_items > 0) {
    const int4 first = reinterpret_cast<const int4*>(items)[0];
    load_D(Dn, first.x, (first.y >> 8) & 0xFF);
  }
  // descriptors travel as int4 {p, meta, next_p, next2_p} / {node, acc_off, rowlen, row_start}
  const int4* items4 = reinterpret_cast<const int4*>(items);
  const int4* nodes4 = reinterpret_cast<const int4*>(nodes);
  int4        un4    = make_int4(0, 0, -1, -1);
  if (n_items > 0) un4 = items4[0];
  for (int u = 0; u < n_items; u++) {
    // the next item's descriptor and this item's row are read now and used much later: no shared-memory round trip sits
    // between two stages
    const int4 nxt4 = items4[u + 1 < n_items ? u + 1 : u];
    struct { int p, meta, next_p, next2_p; } un{un4.x, un4.y, un4.z, un4.w};
    const int  I     = un.meta & 0xFF;
    const int  bg    = (un.meta >> 8) & 0xFF;
    const int4 nd4 = nodes4[(un.meta >> 16) & 0xFF];
    struct { int acc_off, rowlen; } nd{nd4.y, nd4.z};
    const int  b     = bg * C::CPW + cl;
    const bool active = cl < C::CPW && b < ck;
    const bool act1   = active && jz1 < nk;
    const bool act2   = active && col < nk * nk;
    const long long key = (long long)un.p * 256 + bg;
    if (key != cur_key) {  // uniform across the warp
      cur_key = key;
      // Dn was filled for exactly this (element, block) one key ago (or by the prologue): take it over, then send the
      // loads for the next key on their way -- they are not waited for until the warp gets there -- and pull the element
      // after that into L2
      SB_UNROLL
      for (int qz = 0; qz < q; qz++) {
        SB_UNROLL
        for (int i = 0; i < C::NE; i++) D[qz][i] = Dn[qz][i];
      }
      if (un.next_p >= 0) load_D(Dn, un.next_p, (un.meta >> 24) & 0xFF);
      if (un.next2_p >= 0) prefetch_element(un.next2_p);
    }
    // positions of this lane's stage-x outputs in the row (issued early: consumed after both exchanges)
    unsigned short pj[nk];
    if (act2) {
      const unsigned short* pp = P.pos + ((long long)un.p * C::NDT + I) * C::NDK + (jz2 * nk + jy2) * nk;
      SB_UNROLL
      for (int jx = 0; jx < nk; jx++) pj[jx] = pp[jx];
    }

    // ---- stage z: contract the derivative with the test function at the column's points, then with B/G(qz, jz) ----
    if (active) {
      const int    ix = I % nt, iy = (I / nt) % nt, iz = I / (nt * nt);
      const double bx = Bt[qx0 * nt + ix], gx = Gt[qx0 * nt + ix];
      const double by = Bt[qy0 * nt + iy], gy = Gt[qy0 * nt + iy];
      const double bxy = bx * by * wxy, g0xy = gx * by * wxy, g1xy = bx * gy * wxy;
      double p0[nk], p1[nk], p2[nk];
      SB_UNROLL
      for (int j = 0; j < nk; j++) p0[j] = p1[j] = p2[j] = 0.0;
      SB_UNROLL
      for (int qz = 0; qz < q; qz++) {
        const double bz = Btw[qz * nt + iz], gz = Gtw[qz * nt + iz];
        const double phi = bxy * bz;
        const double g[3] = {g0xy * bz, g1xy * bz, bxy * gz};
        double src = 0.0, fl[3] = {0.0, 0.0, 0.0};
        int    i = 0;
        if constexpr (C::d00) src += D[qz][i++] * phi;
        if constexpr (C::d01) {
          SB_UNROLL
          for (int l = 0; l < 3; l++) fl[l] += D[qz][i++] * phi;
        }
        if constexpr (C::d10) {
          SB_UNROLL
          for (int k = 0; k < 3; k++) src += D[qz][i++] * g[k];
        }
        if constexpr (C::d11) {
          SB_UNROLL
          for (int k = 0; k < 3; k++) {
            SB_UNROLL
            for (int l = 0; l < 3; l++) fl[l] += D[qz][i++] * g[k];
          }
        }
        SB_UNROLL
        for (int j = 0; j < nk; j++) {
          const double Bz = P.trial_basis.B[qz * nk + j];
          if constexpr (C::has_source) p0[j] += Bz * src;
          if constexpr (C::has_flux) {
            p0[j] += P.trial_basis.G[qz * nk + j] * fl[2];
            p1[j] += Bz * fl[0];
            p2[j] += Bz * fl[1];
          }
        }
      }
      SB_UNROLL
      for (int j = 0; j < nk; j++) {
        Ps[(0 * nk + j) * SP + lane] = p0[j];
        if constexpr (C::has_flux) {
          Ps[(1 * nk + j) * SP + lane] = p1[j];
          Ps[(2 * nk + j) * SP + lane] = p2[j];
        }
      }
    }
    __syncwarp();
    // ---- stage y ----
    if (act1) {
      double q0[nk], q1[nk];
      SB_UNROLL
      for (int j = 0; j < nk; j++) q0[j] = q1[j] = 0.0;
      SB_UNROLL
      for (int qy = 0; qy < q; qy++) {
        const int    at = cl * QQ + qy * q + qx1;
        const double v0 = Ps[(0 * nk + jz1) * SP + at];
        double       v1 = 0.0, v2 = 0.0;
        if constexpr (C::has_flux) {
          v1 = Ps[(1 * nk + jz1) * SP + at];
          v2 = Ps[(2 * nk + jz1) * SP + at];
        }
        SB_UNROLL
        for (int j = 0; j < nk; j++) {
          const double By = P.trial_basis.B[qy * nk + j];
          q0[j] += By * v0;
          if constexpr (C::has_flux) {
            q0[j] += P.trial_basis.G[qy * nk + j] * v2;
            q1[j] += By * v1;
          }
        }
      }
      SB_UNROLL
      for (int j = 0; j < nk; j++) {
        Qs[(0 * nk + j) * SQ + lane] = q0[j];
        if constexpr (C::has_flux) Qs[(1 * nk + j) * SQ + lane] = q1[j];
      }
    }
    __syncwarp();
    // ---- stage x, and the row (a, I) takes the contribution of this element ----
    if (act2) {
      double* rowa = acc + nd.acc_off;
      double r[nk];
      SB_UNROLL
      for (int j = 0; j < nk; j++) r[j] = 0.0;
      SB_UNROLL
      for (int qx = 0; qx < q; qx++) {
        const int    at = cl * QQ + jz2 * q + qx;
        const double v0 = Qs[(0 * nk + jy2) * SQ + at];
        double       v1 = 0.0;
        if constexpr (C::has_flux) v1 = Qs[(1 * nk + jy2) * SQ + at];
        SB_UNROLL
        for (int j = 0; j < nk; j++) {
          r[j] += P.trial_basis.B[qx * nk + j] * v0;
          if constexpr (C::has_flux) r[j] += P.trial_basis.G[qx * nk + j] * v1;
        }
      }
      SB_UNROLL
      for (int jx = 0; jx < nk; jx++) {
        const int at = P.col_byvdim ? int(pj[jx]) * ck + b : b * nd.rowlen + int(pj[jx]);
        rowa[at] += r[jx];
      }
    }
    un4 = nxt4;
    // the next item's stage z writes Ps (last read before the second __syncwarp) and its stage y writes Qs after its own
    // first one; an entry of the accumulator is touched by one lane per item
  }
  __syncwarp();

  // ---- whole rows to `values` ----
  for (int s = 0; s < grp.node_end - grp.node_begin; s++) {
    const AsmNode   nd  = nodes[s];
    const int       len = nd.rowlen * ck;
    double*         dst = P.values + nd.row_start;
    const double*   src = acc + nd.acc_off;
    if (P.accumulate) {
      for (int j = lane; j < len; j += 32) dst[j] += src[j];
    } else {
      for (int j = lane; j < len; j += 32) dst[j] = src[j];
    }
  }
}

}  // namespace b200
}  // namespace serac
