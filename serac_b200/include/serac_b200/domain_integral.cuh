// serac_b200: fused domain-integral kernels for tensor-product H1 elements (quadrilateral / hexahedron).
//
// What the reference does per element on the host, one kernel here does per CTA of EPB elements:
//   * L->E gather            src/serac/numerics/functional/element_restriction.cpp:443-454   (fused: reads L through the
//                                                                                              restriction table)
//   * interpolate            src/serac/numerics/functional/detail/hexahedron_H1.inl:169-228  (sum-factorised, smem staged)
//   * parent_to_physical     src/serac/numerics/functional/finite_element.hpp:246-273
//   * q-function (+ state)   src/serac/numerics/functional/domain_integral_kernels.hpp:105-150
//   * physical_to_parent     src/serac/numerics/functional/finite_element.hpp:287-322
//   * store d(qf)/d(arg)     src/serac/numerics/functional/domain_integral_kernels.hpp:206-213
//   * integrate              src/serac/numerics/functional/detail/hexahedron_H1.inl:230-280  (sum-factorised, smem staged)
// and the two derivative consumers
//   * action_of_gradient     src/serac/numerics/functional/domain_integral_kernels.hpp:277-304
//   * element_gradient       src/serac/numerics/functional/domain_integral_kernels.hpp:327-361
//
// Thread mapping: one thread per quadrature point, EPB elements per CTA (EPB * q^dim ~ 128 threads).  All
// per-quadrature-point streams (stored derivatives, geometric factors) are laid out [tile][entry][lane] so that a
// warp reads consecutive doubles.  The element residual is written as an E-vector block; the deterministic E->L
// scatter is a separate kernel (restriction.cu).
#pragma once

#include <cuda_runtime.h>

#include "serac_b200/finite_element.hpp"
#include "serac_b200/physics_functors.hpp"
#include "serac_b200/tensor.hpp"

namespace serac {
namespace b200 {

constexpr int MAX_TRIALS = 6;

/// how many elements one CTA processes
SB_HD int elements_per_block(int dim, int q)
{
  // q = 3: two elements (54 threads, 35 KB of stored derivatives per bulk copy) and four CTAs per SM in the streaming JVP
  // kernel instead of four elements and two CTAs: the same bytes in flight per SM in finer grains (109 -> 105 us on
  // config A).  Possible since the derivative layout of these elements no longer depends on the CTA size.
  if (dim == 3) return (q <= 1) ? 64 : (q == 2) ? 16 : 2;
  if (dim == 2) return (q <= 1) ? 128 : (q == 2) ? 32 : (q == 3) ? 14 : 8;
  return 1;
}

/// geometry-aware version: simplices have no tensor structure, one thread per quadrature point, ~128 threads per CTA
SB_HD int elements_per_block(Geometry g, int q)
{
  if (!uses_dense_table(g)) return elements_per_block(dimension_of(g), q);
  const int nq  = num_quadrature_points(g, q);
  const int epb = 128 / nq;
  return epb > 32 ? 32 : (epb < 1 ? 1 : epb);
}

/// Stored-derivative layout: [tile][entry][tile elements x QPE lanes].  Elements with a warp's worth of quadrature points
/// (hexahedra with q >= 3) are their own tile, i.e. [element][entry][QPE]: the evaluation kernel's consecutive stores then
/// fall on consecutive addresses (config A: 159 -> 150 us), and any kernel may stream any number of elements per CTA.
/// Smaller elements keep EPB elements per tile so that a warp's access to one entry stays one contiguous run.
SB_HD int dq_tile_elements(bool simplex, int qpe, int epb) { return (!simplex && qpe >= 27) ? 1 : epb; }

/// one trial / test field as the kernels see it
struct FieldArg {
  const double* L;            ///< L-vector (device)
  const int*    table;        ///< [elements of this geometry][ndof] lexicographic scalar node ids
  long long     node_stride;  ///< byVDIM: components ; byNODES: 1      (element_restriction.cpp:425-432)
  long long     comp_stride;  ///< byVDIM: 1          ; byNODES: num_nodes
};

template <int NT>
struct EvalParams {
  FieldArg      trial[NT > 0 ? NT : 1];
  BasisTable1D  trial_basis[NT > 0 ? NT : 1];
  BasisTable1D  test_basis;
  const double* simplex_tab[NT + 1];  ///< simplex geometries: device SimplexTable::tab of every trial space, then test
  const double* X_q;           ///< [e][i][q]       (geometric_factors.cpp:27,48)
  const double* J_q;           ///< [e][j][i][q] = dX_i/dxi_j
  const int*    elements;      ///< domain element ids (index into the restriction tables / E-vector)
  int           num_elements;  ///< END of the element range of this launch (exclusive): positions in the domain's list
  int           first_element; ///< start of the range (0 for a whole-domain launch); the grid covers [first, num_elements)
  double*       r_E;           ///< [element][comp][node] E-vector block of the test space
  double*       dqf;           ///< stored q-function derivatives, [tile][entry][lane]
  void*         qdata;         ///< [e][q] AoS of the state type (quadrature_data.hpp:127-131)
  int           update_state;
  double        t;
  int           prefetch_distance;  ///< CTAs resident on the device: tile blockIdx + distance is prefetched into L2 (0: off)
  int           node_major;         ///< r_E layout: 0 = [element][comp][node] (reference), 1 = [element][node][comp]
};

/// position of (element, component a, node) in an E-vector block of ct components and ndof nodes per element
__device__ __forceinline__ long long e_index(int node_major, long long elem, int a, int node, int ct, int ndof)
{
  return node_major ? (elem * ndof + node) * ct + a : (elem * ct + a) * ndof + node;
}

/// one 128-byte line into L2 (no register, no scoreboard: the CTA that runs `prefetch_distance` tiles later finds its
/// streaming inputs -- restriction-table rows, X_q, J_q -- in L2 instead of paying the HBM latency twice in a row)
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int THREADS>
__device__ __forceinline__ void prefetch_range_l2(const void* base, long long bytes)
{
  const char* b = static_cast<const char*>(base);
  for (long long off = (long long)threadIdx.x * 128; off < bytes; off += (long long)THREADS * 128) prefetch_l2(b + off);
}

struct JvpParams {
  FieldArg      trial;
  BasisTable1D  trial_basis;
  BasisTable1D  test_basis;
  const double* simplex_trial_tab;  ///< simplex geometries only
  const double* simplex_test_tab;
  const int*    elements;
  int           num_elements;
  const double* dqf;
  double*       r_E;
  int           node_major;  ///< r_E layout (see EvalParams)
};

struct ElemGradParams {
  BasisTable1D  trial_basis;
  BasisTable1D  test_basis;
  const double* simplex_trial_tab;  ///< simplex geometries only
  const double* simplex_test_tab;
  const int*    elements;
  int           num_elements;
  const double* dqf;
  double*       K_e;  ///< [element][trial_comp*ntrial + J][test_comp*ntest + I]   (functional.hpp:599-601,632-636)
  int           node_major;  ///< 1: [element][I][test_comp][J][trial_comp] instead
  int           overwrite;   ///< 1: store instead of accumulate
};

/// position of K_e(element; trial comp b, trial node J; test comp a, test node I)
__device__ __forceinline__ long long k_index(int node_major, long long elem, int b, int J, int a, int I, int ck, int ndk,
                                             int ct, int ndt)
{
  return node_major ? (((elem * ndt + I) * ct + a) * ndk + J) * ck + b
                    : (elem * (ck * ndk) + (b * ndk + J)) * (long long)(ct * ndt) + a * ndt + I;
}

// ------------------------------------------------------------------------------------------------
// component accessors that hide the c == 1 special cases of the q-function argument types
// (domain_integral_kernels.hpp:33-65)
// ------------------------------------------------------------------------------------------------
template <typename Space, int dim, typename T = double>
struct qf_arg {
  static constexpr int c = Space::components;
  using value_t          = std::conditional_t<c == 1, T, tensor<T, c>>;
  using grad_t           = std::conditional_t<c == 1, tensor<T, dim>, tensor<T, c, dim>>;
  using type             = tuple<value_t, grad_t>;
};

template <typename T, typename = std::enable_if_t<is_scalar_v<T>>>
SB_HD T& vcomp(T& v, int)
{
  return v;
}
template <typename T, typename = std::enable_if_t<is_scalar_v<T>>>
SB_HD const T& vcomp(const T& v, int)
{
  return v;
}
template <typename T, int c>
SB_HD T& vcomp(tensor<T, c>& v, int a)
{
  return v[a];
}
template <typename T, int c>
SB_HD const T& vcomp(const tensor<T, c>& v, int a)
{
  return v[a];
}
template <typename T, int dim>
SB_HD T& gcomp(tensor<T, dim>& g, int, int j)
{
  return g[j];
}
template <typename T, int dim>
SB_HD const T& gcomp(const tensor<T, dim>& g, int, int j)
{
  return g[j];
}
template <typename T, int c, int dim>
SB_HD T& gcomp(tensor<T, c, dim>& g, int a, int j)
{
  return g[a][j];
}
template <typename T, int c, int dim>
SB_HD const T& gcomp(const tensor<T, c, dim>& g, int a, int j)
{
  return g[a][j];
}

template <typename T>
struct scalar_type_of {
  using type = T;
};
template <typename T, int... n>
struct scalar_type_of<tensor<T, n...>> {
  using type = T;
};
template <typename T>
inline constexpr bool carries_derivative_v = is_dual_number<typename scalar_type_of<T>::type>::value;

/// partial k of a (possibly plain) scalar
SB_HD double partial(double, int) { return 0.0; }
template <int N>
SB_HD double partial(const dual<N>& x, int k)
{
  return x.gradient[k];
}

// ------------------------------------------------------------------------------------------------
// sum-factorised interpolation / integration staged through shared memory.
// All threads of the CTA must call these (they contain __syncthreads()).
// Thread (qz,qy,qx) of an element owns quadrature point Q = (qz*q + qy)*q + qx   (hexahedron_H1.inl:155,256).
// ------------------------------------------------------------------------------------------------
template <int dim, int n, int q, int c>
struct TensorProduct;

template <int n, int q, int c>
struct TensorProduct<3, n, q, c> {
  static constexpr int ndof          = n * n * n;
  static constexpr int interp_scratch = c * ndof + 2 * c * n * n * q + 3 * c * n * q * q;
  static constexpr int integ_scratch  = 4 * c * q * q * q + 3 * c * n * q * q + 2 * c * n * n * q;

  /// u: smem [c][n][n][n] already filled; s: this element's scratch (>= interp_scratch doubles, u lives at s[0])
  template <bool NEED_VALUE, bool NEED_GRAD>
  static __device__ __forceinline__ void interpolate(const double* __restrict__ B, const double* __restrict__ G,
                                                     double* s, int qx, int qy, int qz, double (&val)[c],
                                                     double (&grad)[c][3])
  {
    const double* u  = s;
    double*       A1 = s + c * ndof;               // [2][c][n][n][q]
    double*       A2 = A1 + 2 * c * n * n * q;     // [3][c][n][q][q]
    if (qz < n && qy < n) {
      SB_UNROLL
      for (int a = 0; a < c; a++) {
        double sB = 0.0, sG = 0.0;
        SB_UNROLL
        for (int dx = 0; dx < n; dx++) {
          const double uu = u[((a * n + qz) * n + qy) * n + dx];
          sB += B[qx * n + dx] * uu;
          if (NEED_GRAD) sG += G[qx * n + dx] * uu;
        }
        A1[(((0 * c + a) * n + qz) * n + qy) * q + qx] = sB;
        if (NEED_GRAD) A1[(((1 * c + a) * n + qz) * n + qy) * q + qx] = sG;
      }
    }
    __syncthreads();
    if (qz < n) {
      SB_UNROLL
      for (int a = 0; a < c; a++) {
        double sBB = 0.0, sGB = 0.0, sBG = 0.0;
        SB_UNROLL
        for (int dy = 0; dy < n; dy++) {
          const double b  = A1[(((0 * c + a) * n + qz) * n + dy) * q + qx];
          const double By = B[qy * n + dy];
          sBB += By * b;
          if (NEED_GRAD) {
            const double g = A1[(((1 * c + a) * n + qz) * n + dy) * q + qx];
            sGB += By * g;
            sBG += G[qy * n + dy] * b;
          }
        }
        A2[(((0 * c + a) * n + qz) * q + qy) * q + qx] = sBB;
        if (NEED_GRAD) {
          A2[(((1 * c + a) * n + qz) * q + qy) * q + qx] = sGB;
          A2[(((2 * c + a) * n + qz) * q + qy) * q + qx] = sBG;
        }
      }
    }
    __syncthreads();
    SB_UNROLL
    for (int a = 0; a < c; a++) {
      double v = 0.0, gx = 0.0, gy = 0.0, gz = 0.0;
      SB_UNROLL
      for (int dz = 0; dz < n; dz++) {
        const double bb = A2[(((0 * c + a) * n + dz) * q + qy) * q + qx];
        const double Bz = B[qz * n + dz];
        if (NEED_VALUE) v += Bz * bb;
        if (NEED_GRAD) {
          const double gb = A2[(((1 * c + a) * n + dz) * q + qy) * q + qx];
          const double bg = A2[(((2 * c + a) * n + dz) * q + qy) * q + qx];
          gx += Bz * gb;
          gy += Bz * bg;
          gz += G[qz * n + dz] * bb;
        }
      }
      val[a]     = v;
      grad[a][0] = gx;
      grad[a][1] = gy;
      grad[a][2] = gz;
    }
  }

  /// integrate weighted source src[c] / parent-space flux[c][3] of this thread's quadrature point against the
  /// test functions; the thread (dz,dy,dx) = (qz,qy,qx) with all three < n returns the dof value in out[c].
  template <bool HAS_SOURCE, bool HAS_FLUX>
  static __device__ __forceinline__ bool integrate(const double* __restrict__ B, const double* __restrict__ G,
                                                   double* s, int qx, int qy, int qz, const double (&src)[c],
                                                   const double (&flux)[c][3], double (&out)[c])
  {
    double* S  = s;                        // [4][c][q][q][q]
    double* P  = S + 4 * c * q * q * q;    // [3][c][n][q][q]
    double* Qs = P + 3 * c * n * q * q;    // [2][c][n][n][q]
    SB_UNROLL
    for (int a = 0; a < c; a++) {
      if (HAS_SOURCE) S[(((0 * c + a) * q + qz) * q + qy) * q + qx] = src[a];
      if (HAS_FLUX) {
        S[(((1 * c + a) * q + qz) * q + qy) * q + qx] = flux[a][0];
        S[(((2 * c + a) * q + qz) * q + qy) * q + qx] = flux[a][1];
        S[(((3 * c + a) * q + qz) * q + qy) * q + qx] = flux[a][2];
      }
    }
    __syncthreads();
    if (qz < n) {  // dz = qz
      SB_UNROLL
      for (int a = 0; a < c; a++) {
        double p0 = 0.0, p1 = 0.0, p2 = 0.0;
        SB_UNROLL
        for (int k = 0; k < q; k++) {
          const double Bz = B[k * n + qz];
          if (HAS_SOURCE) p0 += Bz * S[(((0 * c + a) * q + k) * q + qy) * q + qx];
          if (HAS_FLUX) {
            p0 += G[k * n + qz] * S[(((3 * c + a) * q + k) * q + qy) * q + qx];
            p1 += Bz * S[(((1 * c + a) * q + k) * q + qy) * q + qx];
            p2 += Bz * S[(((2 * c + a) * q + k) * q + qy) * q + qx];
          }
        }
        P[(((0 * c + a) * n + qz) * q + qy) * q + qx] = p0;
        if (HAS_FLUX) {
          P[(((1 * c + a) * n + qz) * q + qy) * q + qx] = p1;
          P[(((2 * c + a) * n + qz) * q + qy) * q + qx] = p2;
        }
      }
    }
    __syncthreads();
    if (qz < n && qy < n) {  // dz = qz, dy = qy
      SB_UNROLL
      for (int a = 0; a < c; a++) {
        double q0 = 0.0, q1 = 0.0;
        SB_UNROLL
        for (int k = 0; k < q; k++) {
          const double By = B[k * n + qy];
          q0 += By * P[(((0 * c + a) * n + qz) * q + k) * q + qx];
          if (HAS_FLUX) {
            q0 += G[k * n + qy] * P[(((2 * c + a) * n + qz) * q + k) * q + qx];
            q1 += By * P[(((1 * c + a) * n + qz) * q + k) * q + qx];
          }
        }
        Qs[(((0 * c + a) * n + qz) * n + qy) * q + qx] = q0;
        if (HAS_FLUX) Qs[(((1 * c + a) * n + qz) * n + qy) * q + qx] = q1;
      }
    }
    __syncthreads();
    const bool owner = (qz < n && qy < n && qx < n);
    if (owner) {
      SB_UNROLL
      for (int a = 0; a < c; a++) {
        double r = 0.0;
        SB_UNROLL
        for (int k = 0; k < q; k++) {
          r += B[k * n + qx] * Qs[(((0 * c + a) * n + qz) * n + qy) * q + k];
          if (HAS_FLUX) r += G[k * n + qx] * Qs[(((1 * c + a) * n + qz) * n + qy) * q + k];
        }
        out[a] = r;
      }
    }
    return owner;
  }
  /// integrate() with the test space's tables behind an accessor (SmemTables / ParamTables), the columns a stage needs
  /// hoisted into registers; same products in the same order, hence the same bits
  template <bool HAS_SOURCE, bool HAS_FLUX, typename Tab>
  static __device__ __forceinline__ bool integrate_t(const Tab& tab, double* s, int qx, int qy, int qz,
                                                     const double (&src)[c], const double (&flux)[c][3], double (&out)[c])
  {
    double* S  = s;                        // [4][c][q][q][q]
    double* P  = S + 4 * c * q * q * q;    // [3][c][n][q][q]
    double* Qs = P + 3 * c * n * q * q;    // [2][c][n][n][q]
    SB_UNROLL
    for (int a = 0; a < c; a++) {
      if (HAS_SOURCE) S[(((0 * c + a) * q + qz) * q + qy) * q + qx] = src[a];
      if (HAS_FLUX) {
        S[(((1 * c + a) * q + qz) * q + qy) * q + qx] = flux[a][0];
        S[(((2 * c + a) * q + qz) * q + qy) * q + qx] = flux[a][1];
        S[(((3 * c + a) * q + qz) * q + qy) * q + qx] = flux[a][2];
      }
    }
    __syncthreads();
    if (qz < n) {  // dz = qz
      double bz[q], gz[q];
      tab.template col<n, q>(qz, bz, gz);
      SB_UNROLL
      for (int a = 0; a < c; a++) {
        double p0 = 0.0, p1 = 0.0, p2 = 0.0;
        SB_UNROLL
        for (int k = 0; k < q; k++) {
          if (HAS_SOURCE) p0 += bz[k] * S[(((0 * c + a) * q + k) * q + qy) * q + qx];
          if (HAS_FLUX) {
            p0 += gz[k] * S[(((3 * c + a) * q + k) * q + qy) * q + qx];
            p1 += bz[k] * S[(((1 * c + a) * q + k) * q + qy) * q + qx];
            p2 += bz[k] * S[(((2 * c + a) * q + k) * q + qy) * q + qx];
          }
        }
        P[(((0 * c + a) * n + qz) * q + qy) * q + qx] = p0;
        if (HAS_FLUX) {
          P[(((1 * c + a) * n + qz) * q + qy) * q + qx] = p1;
          P[(((2 * c + a) * n + qz) * q + qy) * q + qx] = p2;
        }
      }
    }
    __syncthreads();
    if (qz < n && qy < n) {  // dz = qz, dy = qy
      double by[q], gy[q];
      tab.template col<n, q>(qy, by, gy);
      SB_UNROLL
      for (int a = 0; a < c; a++) {
        double q0 = 0.0, q1 = 0.0;
        SB_UNROLL
        for (int k = 0; k < q; k++) {
          q0 += by[k] * P[(((0 * c + a) * n + qz) * q + k) * q + qx];
          if (HAS_FLUX) {
            q0 += gy[k] * P[(((2 * c + a) * n + qz) * q + k) * q + qx];
            q1 += by[k] * P[(((1 * c + a) * n + qz) * q + k) * q + qx];
          }
        }
        Qs[(((0 * c + a) * n + qz) * n + qy) * q + qx] = q0;
        if (HAS_FLUX) Qs[(((1 * c + a) * n + qz) * n + qy) * q + qx] = q1;
      }
    }
    __syncthreads();
    const bool owner = (qz < n && qy < n && qx < n);
    if (owner) {
      double bx[q], gx[q];
      tab.template col<n, q>(qx, bx, gx);
      SB_UNROLL
      for (int a = 0; a < c; a++) {
        double r = 0.0;
        SB_UNROLL
        for (int k = 0; k < q; k++) {
          r += bx[k] * Qs[(((0 * c + a) * n + qz) * n + qy) * q + k];
          if (HAS_FLUX) r += gx[k] * Qs[(((1 * c + a) * n + qz) * n + qy) * q + k];
        }
        out[a] = r;
      }
    }
    return owner;
  }
};

template <int n, int q, int c>
struct TensorProduct<2, n, q, c> {
  static constexpr int ndof           = n * n;
  static constexpr int interp_scratch = c * ndof + 2 * c * n * q;
  static constexpr int integ_scratch  = 3 * c * q * q + 2 * c * n * q;

  template <bool NEED_VALUE, bool NEED_GRAD>
  static __device__ __forceinline__ void interpolate(const double* __restrict__ B, const double* __restrict__ G,
                                                     double* s, int qx, int qy, int, double (&val)[c],
                                                     double (&grad)[c][2])
  {
    const double* u  = s;
    double*       A1 = s + c * ndof;  // [2][c][n][q]
    if (qy < n) {
      SB_UNROLL
      for (int a = 0; a < c; a++) {
        double sB = 0.0, sG = 0.0;
        SB_UNROLL
        for (int dx = 0; dx < n; dx++) {
          const double uu = u[(a * n + qy) * n + dx];
          sB += B[qx * n + dx] * uu;
          if (NEED_GRAD) sG += G[qx * n + dx] * uu;
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
        const double b  = A1[((0 * c + a) * n + dy) * q + qx];
        const double By = B[qy * n + dy];
        if (NEED_VALUE) v += By * b;
        if (NEED_GRAD) {
          gx += By * A1[((1 * c + a) * n + dy) * q + qx];
          gy += G[qy * n + dy] * b;
        }
      }
      val[a]     = v;
      grad[a][0] = gx;
      grad[a][1] = gy;
    }
  }

  template <bool HAS_SOURCE, bool HAS_FLUX>
  static __device__ __forceinline__ bool integrate(const double* __restrict__ B, const double* __restrict__ G,
                                                   double* s, int qx, int qy, int, const double (&src)[c],
                                                   const double (&flux)[c][2], double (&out)[c])
  {
    double* S = s;                  // [3][c][q][q]
    double* P = S + 3 * c * q * q;  // [2][c][n][q]
    SB_UNROLL
    for (int a = 0; a < c; a++) {
      if (HAS_SOURCE) S[((0 * c + a) * q + qy) * q + qx] = src[a];
      if (HAS_FLUX) {
        S[((1 * c + a) * q + qy) * q + qx] = flux[a][0];
        S[((2 * c + a) * q + qy) * q + qx] = flux[a][1];
      }
    }
    __syncthreads();
    if (qy < n) {  // dy = qy
      SB_UNROLL
      for (int a = 0; a < c; a++) {
        double p0 = 0.0, p1 = 0.0;
        SB_UNROLL
        for (int k = 0; k < q; k++) {
          const double By = B[k * n + qy];
          if (HAS_SOURCE) p0 += By * S[((0 * c + a) * q + k) * q + qx];
          if (HAS_FLUX) {
            p0 += G[k * n + qy] * S[((2 * c + a) * q + k) * q + qx];
            p1 += By * S[((1 * c + a) * q + k) * q + qx];
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
          r += B[k * n + qx] * P[((0 * c + a) * n + qy) * q + k];
          if (HAS_FLUX) r += G[k * n + qx] * P[((1 * c + a) * n + qy) * q + k];
        }
        out[a] = r;
      }
    }
    return owner;
  }
};

// ------------------------------------------------------------------------------------------------
// simplex elements: dense interpolation / integration (triangle_H1.inl:269-336, tetrahedron_H1.inl:369-429).
// tab = SimplexTable::tab in shared memory: [qpt][node][1 + dim] = {phi, dphi/dxi}.
// ------------------------------------------------------------------------------------------------
template <int dim, int ndof, int nq, int c>
struct SimplexOps {
  static constexpr int integ_scratch = c * (1 + dim) * nq;

  /// u: smem [c][ndof] of this element
  static __device__ __forceinline__ void interpolate(const double* __restrict__ tab, const double* __restrict__ u, int ql,
                                                     double (&val)[c], double (&grad)[c][dim])
  {
    SB_UNROLL
    for (int a = 0; a < c; a++) {
      val[a] = 0.0;
      SB_UNROLL
      for (int j = 0; j < dim; j++) grad[a][j] = 0.0;
    }
    const double* t = tab + ql * ndof * (1 + dim);
    SB_UNROLL
    for (int k = 0; k < ndof; k++) {
      double f[1 + dim];
      SB_UNROLL
      for (int j = 0; j <= dim; j++) f[j] = t[k * (1 + dim) + j];
      SB_UNROLL
      for (int a = 0; a < c; a++) {
        const double uu = u[a * ndof + k];
        val[a] += uu * f[0];
        SB_UNROLL
        for (int j = 0; j < dim; j++) grad[a][j] += uu * f[1 + j];
      }
    }
  }

  /// every thread of the CTA must call this; thread ql < ndof returns the value of node ql in out[c]
  template <bool HAS_SOURCE, bool HAS_FLUX>
  static __device__ __forceinline__ bool integrate(const double* __restrict__ tab, double* S, int ql,
                                                   const double (&src)[c], const double (&flux)[c][dim], double (&out)[c])
  {
    SB_UNROLL
    for (int a = 0; a < c; a++) {
      if (HAS_SOURCE) S[(a * (1 + dim) + 0) * nq + ql] = src[a];
      if (HAS_FLUX) {
        SB_UNROLL
        for (int j = 0; j < dim; j++) S[(a * (1 + dim) + 1 + j) * nq + ql] = flux[a][j];
      }
    }
    __syncthreads();
    const bool owner = ql < ndof;
    if (owner) {
      SB_UNROLL
      for (int a = 0; a < c; a++) out[a] = 0.0;
      for (int Q = 0; Q < nq; Q++) {
        const double* t = tab + (Q * ndof + ql) * (1 + dim);
        SB_UNROLL
        for (int a = 0; a < c; a++) {
          double r = out[a];
          if (HAS_SOURCE) r += S[(a * (1 + dim) + 0) * nq + Q] * t[0];
          if (HAS_FLUX) {
            SB_UNROLL
            for (int j = 0; j < dim; j++) r += S[(a * (1 + dim) + 1 + j) * nq + Q] * t[1 + j];
          }
          out[a] = r;
        }
      }
    }
    return owner;
  }
};

SB_HD int cmax(int a, int b) { return a > b ? a : b; }

// ------------------------------------------------------------------------------------------------
// compile-time description of one integral: geometry, Q, test space, trial spaces
// ------------------------------------------------------------------------------------------------
/// EPB_ = elements per CTA = elements per stored-derivative tile (0: elements_per_block default)
template <Geometry geom, int Q, int EPB_, typename test, typename... trials>
struct SignatureT {
  static constexpr Geometry geometry = geom;
  static constexpr int      dim      = dimension_of(geom);
  static constexpr int      q        = Q;
  static constexpr int      NT       = int(sizeof...(trials));
  static constexpr int      QPE      = num_quadrature_points(geom, Q);  ///< quadrature points per element
  static constexpr bool     simplex  = uses_dense_table(geom);  ///< dense {phi, dphi} tables (simplices, segments)
  static constexpr int      EPB      = (EPB_ % 1000 > 0) ? EPB_ % 1000 : elements_per_block(geom, Q);
  static constexpr int      tune     = EPB_ / 1000;  ///< experiment switches of registry_tune.cu (0 in every product table)
  static constexpr int      threads  = EPB * QPE;
  static constexpr int      DQ_TE    = dq_tile_elements(simplex, QPE, EPB);  ///< elements per stored-derivative tile
  using test_space                   = test;
  template <int i>
  using trial_space = typename serac::detail::type_at<i, trials...>::type;

  static constexpr int test_n    = test::order + 1;
  static constexpr int test_c    = test::components;
  static constexpr int test_ndof = num_nodes(geom, test::order);
  static constexpr int trial_ndof(int i)
  {
    constexpr int v[] = {num_nodes(geom, trials::order)..., 0};
    return v[i];
  }
  static constexpr int trial_n(int i)
  {
    constexpr int v[] = {(trials::order + 1)..., 0};
    return v[i];
  }
  static constexpr int trial_c(int i)
  {
    constexpr int v[] = {trials::components..., 0};
    return v[i];
  }

  static_assert(((trials::order + 1 <= Q) && ...), "quadrature rule must have at least p + 1 points per direction");
  static_assert(test::order + 1 <= Q, "quadrature rule must have at least p + 1 points per direction");

  /// domain integral: the element lives in a space of its own dimension
  static constexpr bool boundary = false;
  static constexpr int  sdim     = dim;
};

template <Geometry geom, int Q, typename test, typename... trials>
using Signature = SignatureT<geom, Q, 0, test, trials...>;

/// boundary integral (boundary_integral_kernels.hpp:159-205): `geom` is the FACE geometry, embedded in sdim = dim + 1
/// space dimensions.  X_q is [e][sdim][q], J_q is [e][dim][sdim][q]; the q-function sees tuple{x, dX/dxi (sdim x dim)},
/// trial gradients stay in parent (xi) coordinates and its result is a source only, weighted by |dX/dxi_1 x dX/dxi_2|.
template <Geometry geom, int Q, int EPB_, typename test, typename... trials>
struct BoundarySignatureT : SignatureT<geom, Q, EPB_, test, trials...> {
  static constexpr bool boundary = true;
  static constexpr int  sdim     = dimension_of(geom) + 1;
};
template <Geometry geom, int Q, typename test, typename... trials>
using BoundarySignature = BoundarySignatureT<geom, Q, 0, test, trials...>;

/// first q-function argument after t: tuple{position, dX/dxi}
template <typename Sig>
using position_t = tuple<tensor<double, Sig::sdim>, tensor<double, Sig::sdim, Sig::dim>>;

/// adapts a boundary q-function (returns the source term alone, boundary_integral_kernels.hpp:111-156) to the
/// tuple{source, flux} convention of the evaluation kernel
template <typename QF>
struct BoundaryIntegrand {
  QF qf;
  template <typename... Args>
  SERAC_HOST_DEVICE auto operator()(double t, const Args&... args) const
  {
    return serac::tuple{qf(t, args...), zero{}};
  }
};

/// elements per CTA of the EVALUATION kernel.  The stored-derivative tiles (and the JVP / element-gradient kernels that
/// stream them) keep Sig::EPB elements per tile; the evaluation kernel may use fewer elements per CTA and then writes
/// its derivatives into the right lanes of the tile its elements belong to.  p = 2 hexes: ONE element = one warp per
/// CTA, 16 CTAs/SM -- block barriers cost nothing and CTAs retire independently (measured: 175 -> 161 us on config A;
/// the JVP stream kernel wants the 4-element tiles: 107 us vs 163 us with 1-element tiles).
#ifndef SB_EVAL_CTA_E_HEX_Q3
#define SB_EVAL_CTA_E_HEX_Q3 1
#endif
#ifndef SB_EVAL_CTA_E_HEX_Q4
#define SB_EVAL_CTA_E_HEX_Q4 1
#endif
template <typename Sig>
constexpr int eval_cta_elements()
{
  if (Sig::geometry == Geometry::CUBE && Sig::q == 3 && !Sig::boundary) {
    return (Sig::EPB % SB_EVAL_CTA_E_HEX_Q3 == 0) ? SB_EVAL_CTA_E_HEX_Q3 : Sig::EPB;
  }
  if (Sig::geometry == Geometry::CUBE && Sig::q == 4 && !Sig::boundary) {  // p = 3: 64 points = two full warps
    return (Sig::EPB % SB_EVAL_CTA_E_HEX_Q4 == 0) ? SB_EVAL_CTA_E_HEX_Q4 : Sig::EPB;
  }
  return Sig::EPB;
}

/// adapts the q-function of a quantity of interest (`Functional<double(trials...)>`, F/functional_qoi.inl: returns the
/// scalar integrand alone) to the tuple{source, flux} convention.  The "test space" is QOI: one node per element whose
/// shape function is 1 (F/detail/qoi.inl), i.e. an order-0 space -- integrate() then is the weighted sum over the
/// element's quadrature points, the E-vector holds one total per element and the scatter adds them up.
template <typename QF>
struct QoiIntegrand {
  QF qf;
  template <typename... Args>
  SERAC_HOST_DEVICE auto operator()(double t, const Args&... args) const
  {
    return serac::tuple{qf(t, args...), zero{}};
  }
};

/// shared-memory plan of the evaluation kernel (doubles):
///   [tables: (NT + 1) x 2 x MAX_Q^2][gathered dofs of every trial field: EPB x c x ndof each][work: EPB x W]
/// The interpolation work areas of the trial fields are disjoint (all fields advance stage by stage together, two
/// barriers in total); the integration scratch of the test space reuses the same work area afterwards.
template <typename Sig>
struct EvalLayout {
  static constexpr int dim = Sig::dim, q = Sig::q, NT = Sig::NT, EPB = eval_cta_elements<Sig>();
  static constexpr int CE      = EPB;                ///< elements per CTA
  static constexpr int threads = CE * Sig::QPE;      ///< threads per CTA
  static constexpr int TILE    = Sig::DQ_TE * Sig::QPE;  ///< lanes of a stored-derivative tile
  static constexpr int QPE = Sig::QPE;
  static constexpr int ndof(int i) { return Sig::trial_ndof(i); }
  /// simplex: offset of space i's {phi, dphi} table (i == NT: test space); the nq weights follow the last table
  static constexpr int simplex_tab_size(int i) { return QPE * (i < NT ? ndof(i) : Sig::test_ndof) * (1 + dim); }
  static constexpr int simplex_tab_off(int i)
  {
    int o = 0;
    for (int j = 0; j < i; j++) o += simplex_tab_size(j);
    return o;
  }
  /// hexahedra: the 1-D tables stay in the kernel's parameter space (ParamTables), no copy in shared memory
  static constexpr bool param_tables = (dim == 3) && !Sig::simplex;
  static constexpr int  tables =
      Sig::simplex ? simplex_tab_off(NT + 1) + QPE : (param_tables ? 0 : (NT + 1) * 2 * MAX_Q * MAX_Q);
  static constexpr int gather_size(int i) { return EPB * Sig::trial_c(i) * ndof(i); }
  static constexpr int gather_off(int i)
  {
    int o = 0;
    for (int j = 0; j < i; j++) o += gather_size(j);
    return o;
  }
  /// per-element interpolation work of field i: A1 (+ A2 in 3D)
  static constexpr int work_size(int i)
  {
    const int n = Sig::trial_n(i), c = Sig::trial_c(i);
    if (Sig::simplex) return 0;  // dense interpolation reads the gathered dofs directly
    return (dim == 3) ? 2 * c * n * n * q + 3 * c * n * q * q : 2 * c * n * q;
  }
  static constexpr int integ_scratch()
  {
    if constexpr (Sig::simplex) {
      return Sig::test_c * (1 + dim) * QPE;
    } else {
      return TensorProduct<dim, Sig::test_n, q, Sig::test_c>::integ_scratch;
    }
  }
  static constexpr int work_off(int i)
  {
    int o = 0;
    for (int j = 0; j < i; j++) o += work_size(j);
    return o;
  }
  static constexpr int W = cmax(work_off(NT), integ_scratch());
  static constexpr int gather_total = gather_off(NT);
  static constexpr int total        = tables + gather_total + EPB * W;
};

template <typename Sig>
constexpr int eval_smem_doubles()
{
  return EvalLayout<Sig>::total;
}

/// the sum-factorised interpolation of one field split into its stages, 1-D table rows hoisted into registers per
/// stage (same staging and summation order as TensorProduct<>::interpolate)
template <int dim, int n, int q, int c>
struct InterpStages;

/// where a kernel finds the 1-D tables: copied into shared memory ...
struct SmemTables {
  const double* B;
  const double* G;
  /// row r of B and G: the n nodal functions (and derivatives) at quadrature point r
  template <int n, int q>
  __device__ __forceinline__ void row(int r, double (&b)[n], double (&g)[n]) const
  {
    SB_UNROLL
    for (int d = 0; d < n; d++) {
      b[d] = B[r * n + d];
      g[d] = G[r * n + d];
    }
  }
  /// column d of B and G: nodal function d at the q quadrature points
  template <int n, int q>
  __device__ __forceinline__ void col(int d, double (&b)[q], double (&g)[q]) const
  {
    SB_UNROLL
    for (int k = 0; k < q; k++) {
      b[k] = B[k * n + d];
      g[k] = G[k * n + d];
    }
  }
};
/// ... or left in the kernel's parameter space (constant bank): a thread picks its row with a chain of selects over
/// compile-time offsets.  No shared-memory traffic: in the hexahedron evaluation kernel 28% of all shared-memory loads
/// were table reads, and the L1 data pipe is that kernel's limiter (profiles/r2q_eval_l1_pipe.txt).
struct ParamTables {
  const BasisTable1D& t;
  template <int n, int q>
  __device__ __forceinline__ void row(int r, double (&b)[n], double (&g)[n]) const
  {
    SB_UNROLL
    for (int d = 0; d < n; d++) {
      double vb = t.B[d], vg = t.G[d];
      SB_UNROLL
      for (int k = 1; k < q; k++) {
        vb = (r == k) ? t.B[k * n + d] : vb;
        vg = (r == k) ? t.G[k * n + d] : vg;
      }
      b[d] = vb;
      g[d] = vg;
    }
  }
  template <int n, int q>
  __device__ __forceinline__ void col(int d, double (&b)[q], double (&g)[q]) const
  {
    SB_UNROLL
    for (int k = 0; k < q; k++) {
      double vb = t.B[k * n], vg = t.G[k * n];
      SB_UNROLL
      for (int j = 1; j < n; j++) {
        vb = (d == j) ? t.B[k * n + j] : vb;
        vg = (d == j) ? t.G[k * n + j] : vg;
      }
      b[k] = vb;
      g[k] = vg;
    }
  }
};

/// NV / NG: does anybody read the interpolated value / gradient?  (qf_usage, physics_functors.hpp)
template <int n, int q, int c>
struct InterpStages<3, n, q, c> {
  template <bool NV = true, bool NG = true, typename Tab>
  static __device__ __forceinline__ void s1(const Tab& tab, const double* __restrict__ u, double* __restrict__ A1, int qx,
                                            int qy, int qz)
  {
    if (qz < n && qy < n) {
      double b[n], g[n];
      tab.template row<n, q>(qx, b, g);
      SB_UNROLL
      for (int a = 0; a < c; a++) {
        double sB = 0.0, sG = 0.0;
        SB_UNROLL
        for (int dx = 0; dx < n; dx++) {
          const double uu = u[((a * n + qz) * n + qy) * n + dx];
          sB += b[dx] * uu;
          if constexpr (NG) sG += g[dx] * uu;
        }
        A1[(((0 * c + a) * n + qz) * n + qy) * q + qx] = sB;
        if constexpr (NG) A1[(((1 * c + a) * n + qz) * n + qy) * q + qx] = sG;
      }
    }
  }
  template <bool NV = true, bool NG = true, typename Tab>
  static __device__ __forceinline__ void s2(const Tab& tab, const double* __restrict__ A1, double* __restrict__ A2, int qx,
                                            int qy, int qz)
  {
    if (qz < n) {
      double b[n], g[n];
      tab.template row<n, q>(qy, b, g);
      SB_UNROLL
      for (int a = 0; a < c; a++) {
        double sBB = 0.0, sGB = 0.0, sBG = 0.0;
        SB_UNROLL
        for (int dy = 0; dy < n; dy++) {
          const double vb = A1[(((0 * c + a) * n + qz) * n + dy) * q + qx];
          sBB += b[dy] * vb;
          if constexpr (NG) {
            const double vg = A1[(((1 * c + a) * n + qz) * n + dy) * q + qx];
            sGB += b[dy] * vg;
            sBG += g[dy] * vb;
          }
        }
        A2[(((0 * c + a) * n + qz) * q + qy) * q + qx] = sBB;
        if constexpr (NG) {
          A2[(((1 * c + a) * n + qz) * q + qy) * q + qx] = sGB;
          A2[(((2 * c + a) * n + qz) * q + qy) * q + qx] = sBG;
        }
      }
    }
  }
  template <bool NV = true, bool NG = true, typename Tab>
  static __device__ __forceinline__ void s3(const Tab& tab, const double* __restrict__ A2, int qx, int qy, int qz,
                                            double (&val)[c], double (&grad)[c][3])
  {
    double b[n], g[n];
    tab.template row<n, q>(qz, b, g);
    SB_UNROLL
    for (int a = 0; a < c; a++) {
      double v = 0.0, gx = 0.0, gy = 0.0, gz = 0.0;
      SB_UNROLL
      for (int dz = 0; dz < n; dz++) {
        const double bb = A2[(((0 * c + a) * n + dz) * q + qy) * q + qx];
        if constexpr (NV) v += b[dz] * bb;
        if constexpr (NG) {
          const double gb = A2[(((1 * c + a) * n + dz) * q + qy) * q + qx];
          const double bg = A2[(((2 * c + a) * n + dz) * q + qy) * q + qx];
          gx += b[dz] * gb;
          gy += b[dz] * bg;
          gz += g[dz] * bb;
        }
      }
      val[a]     = v;
      grad[a][0] = gx;
      grad[a][1] = gy;
      grad[a][2] = gz;
    }
  }
};

template <int n, int q, int c>
struct InterpStages<2, n, q, c> {
  static __device__ __forceinline__ void s1(const double* __restrict__ B, const double* __restrict__ G,
                                            const double* __restrict__ u, double* __restrict__ A1, int qx, int qy, int)
  {
    if (qy < n) {
      double b[n], g[n];
      SB_UNROLL
      for (int d = 0; d < n; d++) {
        b[d] = B[qx * n + d];
        g[d] = G[qx * n + d];
      }
      SB_UNROLL
      for (int a = 0; a < c; a++) {
        double sB = 0.0, sG = 0.0;
        SB_UNROLL
        for (int dx = 0; dx < n; dx++) {
          const double uu = u[(a * n + qy) * n + dx];
          sB += b[dx] * uu;
          sG += g[dx] * uu;
        }
        A1[((0 * c + a) * n + qy) * q + qx] = sB;
        A1[((1 * c + a) * n + qy) * q + qx] = sG;
      }
    }
  }
  /// final stage in 2D
  static __device__ __forceinline__ void s2(const double* __restrict__ B, const double* __restrict__ G,
                                            const double* __restrict__ A1, int qx, int qy, int, double (&val)[c],
                                            double (&grad)[c][2])
  {
    double b[n], g[n];
    SB_UNROLL
    for (int d = 0; d < n; d++) {
      b[d] = B[qy * n + d];
      g[d] = G[qy * n + d];
    }
    SB_UNROLL
    for (int a = 0; a < c; a++) {
      double v = 0.0, gx = 0.0, gy = 0.0;
      SB_UNROLL
      for (int dy = 0; dy < n; dy++) {
        const double vb = A1[((0 * c + a) * n + dy) * q + qx];
        v += b[dy] * vb;
        gx += b[dy] * A1[((1 * c + a) * n + dy) * q + qx];
        gy += g[dy] * vb;
      }
      val[a]     = v;
      grad[a][0] = gx;
      grad[a][1] = gy;
    }
  }
};

// ------------------------------------------------------------------------------------------------
// layout of the stored derivative d(source, flux)/d(value, grad) w.r.t. one trial argument.
// Blocks follow the reference's nested tuple order (domain_integral_kernels.hpp:223-237):
//   d00 = d source / d value   [ct][ck]
//   d01 = d source / d grad    [ct][ck][dim]
//   d10 = d flux   / d value   [ct][dim][ck]
//   d11 = d flux   / d grad    [ct][dim][ck][dim]
// A block is absent when the q-function's output does not carry derivatives in the pass that seeds it -- the same
// structural pruning the reference gets from `zero` gradient types.
// ------------------------------------------------------------------------------------------------
//
// Symmetric storage (an extension; the reference stores every block in full): when the q-function declares that its
// flux tangent has the major symmetry d flux[a][k] / d grad[b][l] = d flux[b][l] / d grad[a][k] (hyperelastic
// materials: the first Piola tangent is the Hessian of the strain energy, and the congruences of the shape-aware
// wrapper and of physical_to_parent keep it), d11 is stored as the upper triangle of the (ct dim) x (ck dim) matrix,
// row-major: N (N + 1) / 2 instead of N^2 doubles per quadrature point (45 instead of 81 for a 3-D vector field).
// Every reader addresses d11 through i11().
// ------------------------------------------------------------------------------------------------
struct DerivativeLayout {
  int  ct, ck, dim;
  bool has[4];
  int  offset[4];
  int  size[4];
  int  entries;
  bool sym;  ///< d11 stored as its upper triangle
  /// is (a, k, b, l) an entry the symmetric layout stores itself (the others are read through their transposes) ?
  SB_HD bool stores11(int a, int k, int b, int l) const { return !sym || (a * dim + k) <= (b * dim + l); }
  /// position of d flux[a][k] / d grad[b][l] among the entries of one quadrature point
  SB_HD int i11(int a, int k, int b, int l) const
  {
    const int i = a * dim + k, j = b * dim + l, N = ck * dim;
    if (!sym) return offset[3] + i * N + j;
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    return offset[3] + lo * N - (lo * (lo - 1)) / 2 + (hi - lo);
  }
};

/// bit 4 of the kernels' block masks: symmetric d11
constexpr int SB_MASK_SYM = 16;

SB_HD DerivativeLayout make_layout(int ct, int ck, int dim, bool d00, bool d01, bool d10, bool d11,
                                             bool sym = false)
{
  DerivativeLayout L{};
  L.ct      = ct;
  L.ck      = ck;
  L.dim     = dim;
  L.has[0]  = d00;
  L.has[1]  = d01;
  L.has[2]  = d10;
  L.has[3]  = d11;
  L.sym     = sym && d11 && ct == ck;
  L.size[0] = ct * ck;
  L.size[1] = ct * ck * dim;
  L.size[2] = ct * dim * ck;
  L.size[3] = L.sym ? (ct * dim * (ct * dim + 1)) / 2 : ct * dim * ck * dim;
  int off   = 0;
  for (int b = 0; b < 4; b++) {
    L.offset[b] = off;
    if (L.has[b]) off += L.size[b];
  }
  L.entries = off;
  return L;
}

/// compile-time check of the symmetric packing: i11 is symmetric under (a,k) <-> (b,l) and maps the upper triangle one to
/// one onto [offset[3], offset[3] + N (N + 1) / 2)
SB_HD bool symmetric_layout_is_consistent(int c, int dim)
{
  const DerivativeLayout L = make_layout(c, c, dim, false, false, false, true, true);
  const int              N = c * dim;
  if (!L.sym || L.entries != N * (N + 1) / 2) return false;
  int next = 0;  // the upper triangle in row-major order must come out as 0, 1, 2, ...
  for (int i = 0; i < N; i++) {
    for (int j = 0; j < N; j++) {
      const int ij = L.i11(i / dim, i % dim, j / dim, j % dim), ji = L.i11(j / dim, j % dim, i / dim, i % dim);
      if (ij != ji || ij < 0 || ij >= L.entries) return false;
      if (j >= i && ij != next++) return false;
      if (L.stores11(i / dim, i % dim, j / dim, j % dim) != (i <= j)) return false;
    }
  }
  return next == L.entries;
}
static_assert(symmetric_layout_is_consistent(3, 3) && symmetric_layout_is_consistent(2, 2) &&
                  symmetric_layout_is_consistent(1, 3) && symmetric_layout_is_consistent(3, 2),
              "DerivativeLayout: symmetric d11 packing");
static_assert(make_layout(3, 3, 3, false, false, false, true, true).entries == 45 &&
                  make_layout(3, 3, 3, false, false, false, true, false).entries == 81 &&
                  !make_layout(3, 1, 3, false, false, false, true, true).sym,
              "DerivativeLayout: 45 of 81 entries for a 3-D vector field; rectangular blocks stay full");

// ---- calling the q-function -------------------------------------------------------------------
template <typename QF, typename Pos, typename State, typename... Args>
SERAC_HOST_DEVICE inline auto call_qf(const QF& qf, double t, const Pos& x, State& state, const Args&... args)
{
  if constexpr (std::is_same_v<State, Nothing>) {
    return qf(t, x, args...);
  } else {
    return qf(t, x, state, args...);
  }
}

template <typename T>
struct output_traits;  // of tuple<source, flux>
template <typename S, typename F>
struct output_traits<tuple<S, F>> {
  static constexpr bool has_source  = !is_zero<S>::value;
  static constexpr bool has_flux    = !is_zero<F>::value;
  static constexpr bool source_dual = carries_derivative_v<S>;
  static constexpr bool flux_dual   = carries_derivative_v<F>;
};

/// argument K as seen by the q-function in a derivative pass
/// PASS == -1: value pass (value components carry dual<ck>, gradient plain)
/// PASS >= 0 : gradient pass: rows [PASS*ROWS, (PASS+1)*ROWS) of the gradient are seeded, every gradient entry is a
///             dual<ROWS*dim> (partial index = local_row*dim + parent direction), value plain
template <typename Space, int dim, int PASS, int ROWS = 1>
struct seeded_arg {
  static constexpr int c = Space::components;
  static constexpr int N = (PASS < 0) ? c : ROWS * dim;
  using value_scalar     = std::conditional_t<(PASS < 0), dual<N>, double>;
  using grad_scalar      = std::conditional_t<(PASS < 0), double, dual<N>>;
  using value_t          = std::conditional_t<c == 1, value_scalar, tensor<value_scalar, c>>;
  using grad_t           = std::conditional_t<c == 1, tensor<grad_scalar, dim>, tensor<grad_scalar, c, dim>>;
  using type             = tuple<value_t, grad_t>;
};

template <typename Sig, int WRT, int PASS, typename QF, typename State, int ROWS, int... I>
auto pass_output_type(std::integer_sequence<int, I...>)
{
  constexpr int dim = Sig::dim;
  using pos_t       = position_t<Sig>;
  if constexpr (std::is_same_v<State, Nothing>) {
    return decltype(std::declval<QF>()(
        double{}, pos_t{},
        std::declval<std::conditional_t<
            I == WRT, typename seeded_arg<typename Sig::template trial_space<I>, dim, PASS, ROWS>::type,
            typename qf_arg<typename Sig::template trial_space<I>, dim>::type>>()...)){};
  } else {
    return decltype(std::declval<QF>()(
        double{}, pos_t{}, std::declval<State&>(),
        std::declval<std::conditional_t<
            I == WRT, typename seeded_arg<typename Sig::template trial_space<I>, dim, PASS, ROWS>::type,
            typename qf_arg<typename Sig::template trial_space<I>, dim>::type>>()...)){};
  }
}

template <typename Sig, typename QF, typename State, int... I>
auto plain_output_type(std::integer_sequence<int, I...>)
{
  constexpr int dim = Sig::dim;
  using pos_t       = position_t<Sig>;
  if constexpr (std::is_same_v<State, Nothing>) {
    return decltype(std::declval<QF>()(
        double{}, pos_t{}, std::declval<typename qf_arg<typename Sig::template trial_space<I>, dim>::type>()...)){};
  } else {
    return decltype(std::declval<QF>()(
        double{}, pos_t{}, std::declval<State&>(),
        std::declval<typename qf_arg<typename Sig::template trial_space<I>, dim>::type>()...)){};
  }
}

/// FAST: use the q-function's own flux tangent (provides_flux_tangent, physics_functors.hpp) instead of dual numbers
template <typename Sig, int WRT, typename QF, typename State, bool FAST = true>
struct DerivativeTraits {
  using seq              = std::make_integer_sequence<int, Sig::NT>;
  using value_pass_out   = decltype(pass_output_type<Sig, WRT, -1, QF, State, 1>(seq{}));
  using grad_pass_out    = decltype(pass_output_type<Sig, WRT, 0, QF, State, 1>(seq{}));
  using trial            = typename Sig::template trial_space<WRT>;
  static constexpr int  ct  = Sig::test_c;
  static constexpr int  ck  = trial::components;
  static constexpr int  dim = Sig::dim;
  static constexpr bool tangent = FAST && !std::is_same_v<State, Nothing> && provides_flux_tangent<QF, WRT>::value;
  static constexpr bool d00 = !tangent && output_traits<value_pass_out>::source_dual;
  static constexpr bool d10 = !tangent && output_traits<value_pass_out>::flux_dual;
  static constexpr bool d01 = !tangent && output_traits<grad_pass_out>::source_dual;
  static constexpr bool d11 = tangent || output_traits<grad_pass_out>::flux_dual;
  static_assert(!tangent || (!output_traits<value_pass_out>::source_dual && !output_traits<value_pass_out>::flux_dual &&
                             !output_traits<grad_pass_out>::source_dual),
                "a flux tangent provider may only be declared for an argument whose gradient alone enters the flux");
  static constexpr bool value_pass_active = d00 || d10;
  static constexpr bool grad_pass_active  = d01 || d11;
  /// upper-triangle storage of d11: only on the tangent path, where the q-function itself vouches for the symmetry
  static constexpr bool sym = tangent && ct == ck && provides_symmetric_flux_tangent<QF, WRT>::value;
  static constexpr DerivativeLayout layout() { return make_layout(ct, ck, dim, d00, d01, d10, d11, sym); }
};

// ---- output component extraction ----------------------------------------------------------------
template <int ct, typename S>
SB_HD auto source_component(const S& s, int a)
{
  if constexpr (is_zero<S>::value) {
    return 0.0;
  } else {
    return vcomp(s, a);
  }
}
template <int ct, int dim, typename F>
SB_HD auto flux_component(const F& f, int a, int j)
{
  if constexpr (is_zero<F>::value) {
    return 0.0;
  } else {
    return gcomp(f, a, j);
  }
}

// ------------------------------------------------------------------------------------------------
// evaluation kernel.  WRT < 0: plain residual.  WRT >= 0: residual + derivative store w.r.t. trial WRT.
// ------------------------------------------------------------------------------------------------
/// build the seeded version of argument `arg` for a derivative pass
template <typename Space, int dim, int PASS, int ROWS>
__device__ __forceinline__ auto seed_argument(const typename qf_arg<Space, dim>::type& arg,
                                              const tensor<double, dim, dim>& invJ)
{
  using S          = seeded_arg<Space, dim, PASS, ROWS>;
  constexpr int c  = Space::components;
  typename S::type out{};
  if constexpr (PASS < 0) {
    SB_UNROLL
    for (int a = 0; a < c; a++) {
      auto& v              = vcomp(get<0>(out), a);
      v                    = make_dual_constant<S::N>(vcomp(get<0>(arg), a));
      v.gradient[a]        = 1.0;
      SB_UNROLL
      for (int j = 0; j < dim; j++) gcomp(get<1>(out), a, j) = gcomp(get<1>(arg), a, j);
    }
  } else {
    SB_UNROLL
    for (int a = 0; a < c; a++) {
      vcomp(get<0>(out), a) = vcomp(get<0>(arg), a);
      SB_UNROLL
      for (int j = 0; j < dim; j++) {
        auto& g = gcomp(get<1>(out), a, j);
        g       = make_dual_constant<S::N>(gcomp(get<1>(arg), a, j));
        // the seed is the PARENT-space gradient (the reference promotes to dual before parent_to_physical,
        // domain_integral_kernels.hpp:181-186), so d(physical grad[a][j]) / d(parent grad[a][k]) = invJ(k, j)
        if (a >= PASS * ROWS && a < (PASS + 1) * ROWS) {
          SB_UNROLL
          for (int k = 0; k < dim; k++) g.gradient[(a - PASS * ROWS) * dim + k] = invJ(k, j);
        }
      }
    }
  }
  return out;
}

template <typename Sig, int WRT, int PASS, int ROWS, typename QF, typename State, typename Pos, typename Inputs, int... I>
__device__ __forceinline__ auto run_pass(const QF& qf, double t, const Pos& pos, State& state, const Inputs& in,
                                         const tensor<double, Sig::dim, Sig::dim>& invJ,
                                         std::integer_sequence<int, I...>)
{
  auto pick = [&](auto idx) {
    constexpr int i = decltype(idx)::value;
    if constexpr (i == WRT) {
      return seed_argument<typename Sig::template trial_space<i>, Sig::dim, PASS, ROWS>(get<i>(in), invJ);
    } else {
      return get<i>(in);
    }
  };
  return call_qf(qf, t, pos, state, pick(std::integral_constant<int, I>{})...);
}

template <typename Sig, typename QF, typename Pos, typename State, typename Inputs, int... I>
__device__ __forceinline__ auto call_qf_tuple(const QF& qf, double t, const Pos& pos, State& st, const Inputs& in,
                                              std::integer_sequence<int, I...>)
{
  return call_qf(qf, t, pos, st, get<I>(in)...);
}

/// ROWS: gradient rows seeded per derivative pass (1 .. ck); MINB: __launch_bounds__ minimum resident CTAs per SM
template <typename Sig, int WRT, typename QF, typename State, int ROWS = 1, int MINB = 1, bool FAST = true>
__global__ void __launch_bounds__(EvalLayout<Sig>::threads, MINB) evaluation_kernel(const __grid_constant__ EvalParams<Sig::NT> P,
                                                                  const __grid_constant__ QF qf)
{
  constexpr int dim = Sig::dim;
  constexpr int q   = Sig::q;
  constexpr int NT  = Sig::NT;
  constexpr int QPE = Sig::QPE;
  constexpr int EPB = EvalLayout<Sig>::CE;  // elements of THIS CTA (the derivative tile has Sig::EPB)
  constexpr int ct  = Sig::test_c;
  using LAY         = EvalLayout<Sig>;
  constexpr int per = LAY::W;
  using seq         = std::make_integer_sequence<int, NT>;

  extern __shared__ double smem[];
  double* tables  = smem;
  double* gathered = smem + LAY::tables;
  double* scratch = gathered + LAY::gather_total;

  if constexpr (Sig::simplex) {
    // {phi, dphi} tables of every space, then the quadrature weights (which trail the test space's table in global memory)
    [&]<int... I>(std::integer_sequence<int, I...>)
    {
      auto copy = [&](auto idx) {
        constexpr int J   = decltype(idx)::value;
        constexpr int off = LAY::simplex_tab_off(J);
        constexpr int sz  = LAY::simplex_tab_size(J) + (J == NT ? QPE : 0);
        const double* src = P.simplex_tab[J];
        for (int i = threadIdx.x; i < sz; i += LAY::threads) tables[off + i] = src[i];
      };
      (copy(std::integral_constant<int, I>{}), ...);
    }
    (std::make_integer_sequence<int, NT + 1>{});
  } else if constexpr (!LAY::param_tables) {
    for (int i = threadIdx.x; i < LAY::tables; i += LAY::threads) {
      const int           which = i / (2 * MAX_Q * MAX_Q);
      const int           rem   = i - which * (2 * MAX_Q * MAX_Q);
      const BasisTable1D& bt    = (which < NT) ? P.trial_basis[which < NT ? which : 0] : P.test_basis;
      tables[i]                 = (rem < MAX_Q * MAX_Q) ? bt.B[rem] : bt.G[rem - MAX_Q * MAX_Q];
    }
  }

  const int el = threadIdx.x / QPE;
  const int ql = threadIdx.x - el * QPE;
  const int qx = ql % q;
  const int qy = (ql / q) % q;
  const int qz = (dim == 3) ? ql / (q * q) : 0;

  const int  e_raw = P.first_element + blockIdx.x * EPB + el;
  const bool valid = e_raw < P.num_elements;
  const int  e     = valid ? e_raw : P.num_elements - 1;
  const int  elem  = P.elements ? P.elements[e] : e;

  constexpr int sdim = Sig::sdim;
  // L2 prefetch of a later tile's streaming inputs (identity element list: the tile's rows are contiguous)
  if (P.prefetch_distance > 0 && !P.elements) {
    const long long pe = P.first_element + ((long long)blockIdx.x + P.prefetch_distance) * EPB;
    if (pe + EPB <= P.num_elements) {
      prefetch_range_l2<LAY::threads>(P.J_q + pe * (dim * sdim * QPE), (long long)EPB * dim * sdim * QPE * 8);
      if constexpr (qf_usage<QF>::position)
        prefetch_range_l2<LAY::threads>(P.X_q + pe * (sdim * QPE), (long long)EPB * sdim * QPE * 8);
      [&]<int... I>(std::integer_sequence<int, I...>)
      {
        (prefetch_range_l2<LAY::threads>(P.trial[I].table + pe * LAY::ndof(I), (long long)EPB * LAY::ndof(I) * 4), ...);
      }
      (seq{});
    }
  }

  // geometric factors of this quadrature point:  Jm(i,j) = dX_i/dxi_j = J_q[e](j,i,q)   (loads issued first)
  tensor<double, sdim>      x{};
  tensor<double, sdim, dim> Jm{};
  SB_UNROLL
  for (int i = 0; i < sdim; i++) {
    if constexpr (qf_usage<QF>::position) x[i] = P.X_q[((long long)e * sdim + i) * QPE + ql];  // (else: nobody reads it)
    SB_UNROLL
    for (int j = 0; j < dim; j++) Jm(i, j) = P.J_q[(((long long)e * dim + j) * sdim + i) * QPE + ql];
  }

  // --- L -> E gather of every trial field at once (element_restriction.cpp:443-454): all loads are issued before the
  //     first one is consumed, so the table -> L latency is paid once per CTA, not once per field ---
  {
    constexpr int NGMAX = [] {
      int m = 1;
      for (int i = 0; i < NT; i++) m = cmax(m, (LAY::gather_size(i) + LAY::threads - 1) / LAY::threads);
      return m;
    }();
    // Three phases over ALL fields, each free of branches (slots past the end are clamped and only their store is
    // predicated): every table load is in flight before the first L-vector load needs its index, every L-vector load
    // before the first store.  Written field by field with `if (i < sz)` around the dependent pair, ptxas kept the pairs
    // in seven reconvergence regions one after the other -- seven table latencies in series, 26% of the kernel's stall
    // samples (profiles/r4_summary.md).
    double    gv[NT > 0 ? NT : 1][NGMAX];
    long long gi[NT > 0 ? NT : 1][NGMAX];
    auto      index_field = [&](auto idx) {
      constexpr int   I  = decltype(idx)::value;
      constexpr int   c  = Sig::trial_c(I);
      constexpr int   nd = LAY::ndof(I);
      constexpr int   NG = (LAY::gather_size(I) + LAY::threads - 1) / LAY::threads;
      constexpr int   sz = LAY::gather_size(I);
      const FieldArg& f  = P.trial[I];
      SB_UNROLL
      for (int j = 0; j < NG; j++) {
        const int i0   = threadIdx.x + j * LAY::threads;
        const int i    = i0 < sz ? i0 : sz - 1;
        const int le   = i / (c * nd);
        const int rem  = i - le * (c * nd);
        const int a    = rem / nd;
        const int node = rem - a * nd;
        int       ge;
        if constexpr (EPB == 1) {
          ge = elem;
        } else {
          int ee = P.first_element + blockIdx.x * EPB + le;
          if (ee >= P.num_elements) ee = P.num_elements - 1;
          ge = P.elements ? P.elements[ee] : ee;
        }
        gi[I][j] = (long long)f.table[(long long)ge * nd + node] * f.node_stride + a * f.comp_stride;
      }
    };
    auto load_field = [&](auto idx) {
      constexpr int   I  = decltype(idx)::value;
      constexpr int   NG = (LAY::gather_size(I) + LAY::threads - 1) / LAY::threads;
      const FieldArg& f  = P.trial[I];
      SB_UNROLL
      for (int j = 0; j < NG; j++) gv[I][j] = f.L[gi[I][j]];
    };
    auto store_field = [&](auto idx) {
      constexpr int I  = decltype(idx)::value;
      constexpr int NG = (LAY::gather_size(I) + LAY::threads - 1) / LAY::threads;
      SB_UNROLL
      for (int j = 0; j < NG; j++) {
        const int i = threadIdx.x + j * LAY::threads;
        constexpr int off = LAY::gather_off(I), sz = LAY::gather_size(I);
        if (i < sz) gathered[off + i] = gv[I][j];
      }
    };
    [&]<int... I>(std::integer_sequence<int, I...>) { (index_field(std::integral_constant<int, I>{}), ...); }
    (seq{});
    [&]<int... I>(std::integer_sequence<int, I...>) { (load_field(std::integral_constant<int, I>{}), ...); }
    (seq{});
    [&]<int... I>(std::integer_sequence<int, I...>) { (store_field(std::integral_constant<int, I>{}), ...); }
    (seq{});
  }
  // domain: parent <-> physical maps.  boundary: trial gradients stay in xi coordinates and the measure is the
  // surface element |dX/dxi_1 x dX/dxi_2| (boundary_integral_kernels.hpp:148-152)
  tensor<double, dim, dim> invJ;
  double                   dv;
  if constexpr (Sig::boundary) {
    invJ = Identity<dim>();
    dv   = norm(cross(Jm));
  } else {
    invJ = inv(Jm);
    dv   = det(Jm);
  }
  __syncthreads();

  // --- interpolate every trial field, stage by stage (hexahedron_H1.inl:169-228), then parent -> physical
  //     (finite_element.hpp:262-264) ---
  auto inputs = [&]<int... I>(std::integer_sequence<int, I...>)
  {
    tuple<typename qf_arg<typename Sig::template trial_space<I>, dim>::type...> in{};
    auto u_of  = [&](auto idx) {
      constexpr int J = decltype(idx)::value;
      constexpr int off = LAY::gather_off(J), sz = Sig::trial_c(J) * LAY::ndof(J);
      return gathered + off + el * sz;
    };
    auto w_of = [&](auto idx) {
      constexpr int off = LAY::work_off(decltype(idx)::value);
      return scratch + el * per + off;
    };
    auto B_of = [&](auto idx) { return tables + (decltype(idx)::value * 2 + 0) * MAX_Q * MAX_Q; };
    auto G_of = [&](auto idx) { return tables + (decltype(idx)::value * 2 + 1) * MAX_Q * MAX_Q; };
    // parent -> physical: grad <- grad . inv(J)
    auto finish = [&](auto idx, const auto& val, const auto& grad) {
      constexpr int J = decltype(idx)::value;
      constexpr int c = Sig::trial_c(J);
      auto&         arg = get<J>(in);
      SB_UNROLL
      for (int a = 0; a < c; a++) {
        vcomp(get<0>(arg), a) = val[a];
        SB_UNROLL
        for (int j = 0; j < dim; j++) {
          double sum = 0.0;
          SB_UNROLL
          for (int k = 0; k < dim; k++) sum += grad[a][k] * invJ(k, j);
          gcomp(get<1>(arg), a, j) = sum;
        }
      }
    };
    if constexpr (Sig::simplex) {
      auto dense = [&](auto idx) {
        constexpr int J = decltype(idx)::value;
        constexpr int c = Sig::trial_c(J);
        double        val[c];
        double        grad[c][dim];
        SimplexOps<dim, LAY::ndof(J), QPE, c>::interpolate(tables + LAY::simplex_tab_off(J), u_of(idx), ql, val, grad);
        finish(idx, val, grad);
      };
      (dense(std::integral_constant<int, I>{}), ...);
    } else if constexpr (dim == 3) {
      // tables straight from the kernel parameters; values / gradients the q-function never reads are not computed
      auto s1 = [&](auto idx) {
        constexpr int J = decltype(idx)::value;
        constexpr int n = Sig::trial_n(J), c = Sig::trial_c(J), use = qf_usage<QF>::argument(J);
        using ST        = InterpStages<3, n, q, c>;
        ST::template s1<(use & USES_VALUE) != 0, (use & USES_GRADIENT) != 0>(ParamTables{P.trial_basis[J]}, u_of(idx),
                                                                             w_of(idx), qx, qy, qz);
      };
      auto s2 = [&](auto idx) {
        constexpr int J = decltype(idx)::value;
        constexpr int n = Sig::trial_n(J), c = Sig::trial_c(J), use = qf_usage<QF>::argument(J);
        using ST        = InterpStages<3, n, q, c>;
        ST::template s2<(use & USES_VALUE) != 0, (use & USES_GRADIENT) != 0>(
            ParamTables{P.trial_basis[J]}, w_of(idx), w_of(idx) + 2 * c * n * n * q, qx, qy, qz);
      };
      auto s3 = [&](auto idx) {
        constexpr int J = decltype(idx)::value;
        constexpr int n = Sig::trial_n(J), c = Sig::trial_c(J), use = qf_usage<QF>::argument(J);
        using ST        = InterpStages<3, n, q, c>;
        double val[c];
        double grad[c][3];
        ST::template s3<(use & USES_VALUE) != 0, (use & USES_GRADIENT) != 0>(
            ParamTables{P.trial_basis[J]}, w_of(idx) + 2 * c * n * n * q, qx, qy, qz, val, grad);
        finish(idx, val, grad);
      };
      (s1(std::integral_constant<int, I>{}), ...);
      __syncthreads();
      (s2(std::integral_constant<int, I>{}), ...);
      __syncthreads();
      (s3(std::integral_constant<int, I>{}), ...);
    } else {
      auto s1 = [&](auto idx) {
        constexpr int J = decltype(idx)::value;
        constexpr int n = Sig::trial_n(J), c = Sig::trial_c(J);
        using ST        = InterpStages<2, n, q, c>;
        ST::s1(B_of(idx), G_of(idx), u_of(idx), w_of(idx), qx, qy, qz);
      };
      auto s2 = [&](auto idx) {
        constexpr int J = decltype(idx)::value;
        constexpr int n = Sig::trial_n(J), c = Sig::trial_c(J);
        using ST        = InterpStages<2, n, q, c>;
        double val[c];
        double grad[c][2];
        ST::s2(B_of(idx), G_of(idx), w_of(idx), qx, qy, qz, val, grad);
        finish(idx, val, grad);
      };
      (s1(std::integral_constant<int, I>{}), ...);
      __syncthreads();
      (s2(std::integral_constant<int, I>{}), ...);
    }
    return in;
  }
  (seq{});

  const auto pos = position_t<Sig>{x, Jm};

  // quadrature weight of this point
  double w;
  if constexpr (Sig::simplex) {
    w = tables[LAY::simplex_tab_off(NT + 1) + ql];
  } else {
    const BasisTable1D& tb = P.test_basis;
    w                      = tb.w[qx] * tb.w[qy];
    if (dim == 3) w *= tb.w[qz];
  }

  // --- state ---
  State  state_old{};
  State* qd = nullptr;
  if constexpr (!std::is_same_v<State, Nothing> && !std::is_empty_v<State>) {
    qd        = reinterpret_cast<State*>(P.qdata) + ((long long)e * QPE + ql);
    state_old = *qd;
  }

  double src[ct];
  double flux[ct][dim];
  bool   have_values = false;
  (void)have_values;

  // physical -> parent on one output tuple, generic in the scalar type   (finite_element.hpp:301-303)
  auto pull_back = [&](const auto& out, auto& s_out, auto& f_out) {
    using out_t = std::decay_t<decltype(out)>;
    SB_UNROLL
    for (int a = 0; a < ct; a++) {
      if constexpr (output_traits<out_t>::has_source) s_out[a] = source_component<ct>(get<0>(out), a) * dv;
      if constexpr (output_traits<out_t>::has_flux) {
        SB_UNROLL
        for (int k = 0; k < dim; k++) {
          auto sum = flux_component<ct, dim>(get<1>(out), a, 0) * invJ(k, 0);
          SB_UNROLL
          for (int j = 1; j < dim; j++) sum = sum + flux_component<ct, dim>(get<1>(out), a, j) * invJ(k, j);
          f_out[a][k] = sum * dv;
        }
      }
    }
  };

  using plain_out_t                = decltype(plain_output_type<Sig, QF, State>(seq{}));
  constexpr bool has_source        = output_traits<plain_out_t>::has_source;
  constexpr bool has_flux          = output_traits<plain_out_t>::has_flux;

  auto take_values = [&](const auto& s_out, const auto& f_out) {
    SB_UNROLL
    for (int a = 0; a < ct; a++) {
      if constexpr (has_source) src[a] = get_value(s_out[a]) * w;
      if constexpr (has_flux) {
        SB_UNROLL
        for (int k = 0; k < dim; k++) flux[a][k] = get_value(f_out[a][k]) * w;
      }
    }
  };

  if constexpr (WRT < 0) {
    State st  = state_old;
    auto  out = call_qf_tuple<Sig>(qf, P.t, pos, st, inputs, seq{});
    double s_out[ct];
    double f_out[ct][dim];
    pull_back(out, s_out, f_out);
    take_values(s_out, f_out);
    if constexpr (!std::is_same_v<State, Nothing> && !std::is_empty_v<State>) {
      if (P.update_state && valid) *qd = st;
    }
  } else {
    using DT                   = DerivativeTraits<Sig, WRT, QF, State, FAST>;
    constexpr DerivativeLayout LY = DT::layout();
    constexpr int              ck = DT::ck;
    // derivative tile of this element and this quadrature point's lane in it
    const long long tile = e_raw / Sig::DQ_TE;
    const int       lane = int(e_raw - tile * Sig::DQ_TE) * QPE + ql;
    double*         dq   = P.dqf + tile * (long long)LY.entries * LAY::TILE + lane;
    bool    first = true;

    if constexpr (DT::tangent) {
      // the q-function differentiates itself (closed-form / material-only AD) and applies the parent maps analytically
      State               st = state_old;
      ParentMap<dim>      map{invJ, dv};
      tensor<double, ct>      s_p{};
      tensor<double, ct, dim> f_p{};
      auto sink = [&](int a, int k, int b, int l, double v) {
        constexpr DerivativeLayout LYs = DT::layout();
        if (LYs.stores11(a, k, b, l) && valid) dq[(long long)LYs.i11(a, k, b, l) * LAY::TILE] = v;
      };
      [&]<int... I>(std::integer_sequence<int, I...>)
      {
        qf.flux_tangent(std::integral_constant<int, WRT>{}, P.t, pos, st, map, s_p, f_p, sink, get<I>(inputs)...);
      }
      (seq{});
      SB_UNROLL
      for (int a = 0; a < ct; a++) {
        if constexpr (has_source) src[a] = s_p[a] * w;
        if constexpr (has_flux) {
          SB_UNROLL
          for (int k = 0; k < dim; k++) flux[a][k] = f_p[a][k] * w;
        }
      }
      if constexpr (!std::is_empty_v<State>) {
        if (P.update_state && valid) *qd = st;
      }
      first = false;
    } else if constexpr (DT::value_pass_active) {
      State st  = state_old;
      auto  out = run_pass<Sig, WRT, -1, 1, QF, State>(qf, P.t, pos, st, inputs, invJ, seq{});
      using out_t = std::decay_t<decltype(out)>;
      using S_t   = std::conditional_t<output_traits<out_t>::source_dual, dual<ck>, double>;
      using F_t   = std::conditional_t<output_traits<out_t>::flux_dual, dual<ck>, double>;
      S_t s_out[ct];
      F_t f_out[ct][dim];
      pull_back(out, s_out, f_out);
      if (first) {
        take_values(s_out, f_out);
        if constexpr (!std::is_same_v<State, Nothing> && !std::is_empty_v<State>) {
          if (P.update_state && valid) *qd = st;
        }
        first = false;
      }
      if (valid) {
        SB_UNROLL
        for (int a = 0; a < ct; a++) {
          if constexpr (LY.has[0]) {
            SB_UNROLL
            for (int b = 0; b < ck; b++) dq[(long long)(LY.offset[0] + a * ck + b) * LAY::TILE] = partial(s_out[a], b);
          }
          if constexpr (LY.has[2]) {
            SB_UNROLL
            for (int k = 0; k < dim; k++) {
              SB_UNROLL
              for (int b = 0; b < ck; b++)
                dq[(long long)(LY.offset[2] + (a * dim + k) * ck + b) * LAY::TILE] = partial(f_out[a][k], b);
            }
          }
        }
      }
    }
    if constexpr (!DT::tangent && DT::grad_pass_active) {
      static_assert(ck % ROWS == 0, "ROWS must divide the number of trial components");
      auto grad_pass = [&](auto pass_c) {
        constexpr int              pass = decltype(pass_c)::value;
        constexpr int              NP   = ROWS * dim;
        constexpr DerivativeLayout LY   = DT::layout();  // (re-declared: captured constexpr locals are not constants)
        State                      st   = state_old;
        auto out    = run_pass<Sig, WRT, pass, ROWS, QF, State>(qf, P.t, pos, st, inputs, invJ, seq{});
        using out_t = std::decay_t<decltype(out)>;
        using S_t   = std::conditional_t<output_traits<out_t>::source_dual, dual<NP>, double>;
        using F_t   = std::conditional_t<output_traits<out_t>::flux_dual, dual<NP>, double>;
        S_t s_out[ct];
        F_t f_out[ct][dim];
        pull_back(out, s_out, f_out);
        if (first) {
          take_values(s_out, f_out);
          if constexpr (!std::is_same_v<State, Nothing> && !std::is_empty_v<State>) {
            if (P.update_state && valid) *qd = st;
          }
          first = false;
        }
        if (valid) {
          SB_UNROLL
          for (int rl = 0; rl < ROWS; rl++) {
            const int row = pass * ROWS + rl;
            SB_UNROLL
            for (int a = 0; a < ct; a++) {
              if constexpr (LY.has[1]) {
                SB_UNROLL
                for (int l = 0; l < dim; l++)
                  dq[(long long)(LY.offset[1] + (a * ck + row) * dim + l) * LAY::TILE] =
                      partial(s_out[a], rl * dim + l);
              }
              if constexpr (LY.has[3]) {
                SB_UNROLL
                for (int k = 0; k < dim; k++) {
                  SB_UNROLL
                  for (int l = 0; l < dim; l++)
                    dq[(long long)(LY.i11(a, k, row, l)) * LAY::TILE] =
                        partial(f_out[a][k], rl * dim + l);
                }
              }
            }
          }
        }
      };
      [&]<int... R>(std::integer_sequence<int, R...>) { (grad_pass(std::integral_constant<int, R>{}), ...); }
      (std::make_integer_sequence<int, ck / ROWS>{});
    }
    if constexpr (!DT::tangent && !DT::value_pass_active && !DT::grad_pass_active) {
      State  st  = state_old;
      auto   out = call_qf_tuple<Sig>(qf, P.t, pos, st, inputs, seq{});
      double s_out[ct];
      double f_out[ct][dim];
      pull_back(out, s_out, f_out);
      take_values(s_out, f_out);
      if constexpr (!std::is_same_v<State, Nothing> && !std::is_empty_v<State>) {
        if (P.update_state && valid) *qd = st;
      }
    }
  }

  // --- integrate against the test space and write the E-vector block ---
  if constexpr (has_source || has_flux) {
    __syncthreads();
    double        out[ct];
    if constexpr (!has_source) {
      SB_UNROLL
      for (int a = 0; a < ct; a++) src[a] = 0.0;
    }
    if constexpr (!has_flux) {
      SB_UNROLL
      for (int a = 0; a < ct; a++) {
        SB_UNROLL
        for (int k = 0; k < dim; k++) flux[a][k] = 0.0;
      }
    }
    bool owner;
    if constexpr (Sig::simplex) {
      owner = SimplexOps<dim, Sig::test_ndof, QPE, ct>::template integrate<has_source, has_flux>(
          tables + LAY::simplex_tab_off(NT), scratch + el * per, ql, src, flux, out);
    } else {
      using TP = TensorProduct<dim, Sig::test_n, q, ct>;
      if constexpr (LAY::param_tables) {
        owner = TP::template integrate_t<has_source, has_flux>(ParamTables{P.test_basis}, scratch + el * per, qx, qy, qz,
                                                               src, flux, out);
      } else {
        const double* B = tables + (NT * 2 + 0) * MAX_Q * MAX_Q;
        const double* G = tables + (NT * 2 + 1) * MAX_Q * MAX_Q;
        owner = TP::template integrate<has_source, has_flux>(B, G, scratch + el * per, qx, qy, qz, src, flux, out);
      }
    }
    if (owner && valid) {
      constexpr int n    = Sig::test_n;
      const int     node = Sig::simplex ? ql : (dim == 3) ? (qz * n + qy) * n + qx : qy * n + qx;
      SB_UNROLL
      for (int a = 0; a < ct; a++) P.r_E[e_index(P.node_major, elem, a, node, ct, Sig::test_ndof)] = out[a];
    }
  }
}


// ------------------------------------------------------------------------------------------------
// Jacobian-vector product with the stored derivatives: dr_E = B^T ( dqf : B du )
// (domain_integral_kernels.hpp:277-304, chain_rule :223-237).  Pure streaming over dqf: this is the HBM-bound kernel.
// MASK bit b set <=> derivative block b (d00,d01,d10,d11) is present.
// ------------------------------------------------------------------------------------------------
template <int dim, int q, int nk, int ck, int nt, int ct, int MASK, int EPB_ = 0>
struct JvpConfig {
  static constexpr int  QPE     = ipow(q, dim);
  static constexpr int  EPB     = (EPB_ > 0) ? EPB_ : elements_per_block(dim, q);
  static constexpr int  threads = EPB * QPE;
  static constexpr int  DQ_TE     = dq_tile_elements(false, QPE, EPB);  ///< elements per stored-derivative tile
  static constexpr int  DQ_STRIDE = DQ_TE * QPE;                          ///< lanes per tile = distance between entries
  static constexpr bool d00 = (MASK & 1) != 0, d01 = (MASK & 2) != 0, d10 = (MASK & 4) != 0, d11 = (MASK & 8) != 0;
  using TPk = TensorProduct<dim, nk, q, ck>;
  using TPt = TensorProduct<dim, nt, q, ct>;
  static constexpr int per     = cmax(TPk::interp_scratch, TPt::integ_scratch);
  static constexpr int smem_doubles = 4 * MAX_Q * MAX_Q + EPB * per;
  static constexpr DerivativeLayout layout() { return make_layout(ct, ck, dim, d00, d01, d10, d11, (MASK & SB_MASK_SYM) != 0); }
};

template <int dim, int q, int nk, int ck, int nt, int ct, int MASK, int EPB_ = 0>
__global__ void __launch_bounds__(JvpConfig<dim, q, nk, ck, nt, ct, MASK, EPB_>::threads)
    jvp_kernel(const __grid_constant__ JvpParams P)
{
  using C                       = JvpConfig<dim, q, nk, ck, nt, ct, MASK, EPB_>;
  constexpr DerivativeLayout LY = C::layout();
  constexpr int QPE = C::QPE, EPB = C::EPB, per = C::per;
  extern __shared__ double smem[];
  double* tables  = smem;
  double* scratch = smem + 4 * MAX_Q * MAX_Q;
  for (int i = threadIdx.x; i < 4 * MAX_Q * MAX_Q; i += C::threads) {
    const int           which = i / (2 * MAX_Q * MAX_Q);
    const int           rem   = i - which * (2 * MAX_Q * MAX_Q);
    const BasisTable1D& bt    = (which == 0) ? P.trial_basis : P.test_basis;
    tables[i]                 = (rem < MAX_Q * MAX_Q) ? bt.B[rem] : bt.G[rem - MAX_Q * MAX_Q];
  }
  const int el = threadIdx.x / QPE;
  const int ql = threadIdx.x - el * QPE;
  const int qx = ql % q;
  const int qy = (ql / q) % q;
  const int qz = (dim == 3) ? ql / (q * q) : 0;

  const int  e_raw = blockIdx.x * EPB + el;
  const bool valid = e_raw < P.num_elements;
  const int  e     = valid ? e_raw : P.num_elements - 1;
  const int  elem  = P.elements ? P.elements[e] : e;

  // issue the derivative loads early: they are independent of the interpolation
  const double* dq = (C::DQ_TE == 1) ? P.dqf + ((long long)blockIdx.x * EPB + el) * LY.entries * QPE + ql
                                     : P.dqf + (long long)blockIdx.x * LY.entries * C::threads + threadIdx.x;

  // gather du
  const FieldArg& f = P.trial;
  for (int i = threadIdx.x; i < EPB * ck * C::TPk::ndof; i += C::threads) {
    const int le   = i / (ck * C::TPk::ndof);
    const int rem  = i - le * (ck * C::TPk::ndof);
    const int a    = rem / C::TPk::ndof;
    const int node = rem - a * C::TPk::ndof;
    int       ee   = blockIdx.x * EPB + le;
    if (ee >= P.num_elements) ee = P.num_elements - 1;
    const int       ge  = P.elements ? P.elements[ee] : ee;
    const long long idx = (long long)f.table[(long long)ge * C::TPk::ndof + node];
    scratch[le * per + a * C::TPk::ndof + node] = f.L[idx * f.node_stride + a * f.comp_stride];
  }
  __syncthreads();
  double v[ck];
  double g[ck][dim];
  C::TPk::template interpolate<(C::d00 || C::d10), (C::d01 || C::d11)>(tables, tables + MAX_Q * MAX_Q,
                                                                        scratch + el * per, qx, qy, qz, v, g);

  const BasisTable1D& tb = P.test_basis;
  double              w  = tb.w[qx] * tb.w[qy];
  if (dim == 3) w *= tb.w[qz];

  constexpr bool has_source = C::d00 || C::d01;
  constexpr bool has_flux   = C::d10 || C::d11;
  double         src[ct];
  double         flux[ct][dim];
  const long long stride = C::DQ_STRIDE;
  SB_UNROLL
  for (int a = 0; a < ct; a++) {
    double s = 0.0;
    if constexpr (C::d00) {
      SB_UNROLL
      for (int b = 0; b < ck; b++) s += __ldcs(dq + (LY.offset[0] + a * ck + b) * stride) * v[b];
    }
    if constexpr (C::d01) {
      SB_UNROLL
      for (int b = 0; b < ck; b++) {
        SB_UNROLL
        for (int l = 0; l < dim; l++) s += __ldcs(dq + (LY.offset[1] + (a * ck + b) * dim + l) * stride) * g[b][l];
      }
    }
    src[a] = s * w;
    SB_UNROLL
    for (int k = 0; k < dim; k++) {
      double fl = 0.0;
      if constexpr (C::d10) {
        SB_UNROLL
        for (int b = 0; b < ck; b++) fl += __ldcs(dq + (LY.offset[2] + (a * dim + k) * ck + b) * stride) * v[b];
      }
      if constexpr (C::d11) {
        SB_UNROLL
        for (int b = 0; b < ck; b++) {
          SB_UNROLL
          for (int l = 0; l < dim; l++)
            fl += __ldcs(dq + (LY.i11(a, k, b, l)) * stride) * g[b][l];
        }
      }
      flux[a][k] = fl * w;
    }
  }

  __syncthreads();
  double     out[ct];
  const bool owner = C::TPt::template integrate<has_source, has_flux>(
      tables + 2 * MAX_Q * MAX_Q, tables + 3 * MAX_Q * MAX_Q, scratch + el * per, qx, qy, qz, src, flux, out);
  if (owner && valid) {
    const int node = (dim == 3) ? (qz * nt + qy) * nt + qx : qy * nt + qx;
    SB_UNROLL
    for (int a = 0; a < ct; a++) P.r_E[e_index(P.node_major, elem, a, node, ct, C::TPt::ndof)] = out[a];
  }
}

// ------------------------------------------------------------------------------------------------
// element stiffness matrices from the stored derivatives (domain_integral_kernels.hpp:327-361,
// batch_apply_shape_fn hexahedron_H1.inl:132-167).  For every trial node J the ck trial components are
// integrated together (ct*ck "components"), K_e is written trial-major (functional.hpp:599-601).
// ------------------------------------------------------------------------------------------------
template <int dim, int q, int nk, int ck, int nt, int ct, int MASK, int EPB_ = 0>
struct ElemGradConfig {
  static constexpr int  QPE     = ipow(q, dim);
  static constexpr int  EPB     = (EPB_ > 0) ? EPB_ : elements_per_block(dim, q);
  static constexpr int  threads = EPB * QPE;
  static constexpr int  DQ_TE     = dq_tile_elements(false, QPE, EPB);  ///< elements per stored-derivative tile
  static constexpr int  DQ_STRIDE = DQ_TE * QPE;                          ///< lanes per tile = distance between entries
  static constexpr bool d00 = (MASK & 1) != 0, d01 = (MASK & 2) != 0, d10 = (MASK & 4) != 0, d11 = (MASK & 8) != 0;
  using TPt = TensorProduct<dim, nt, q, ct * ck>;
  static constexpr int ndof_k = ipow(nk, dim);
  static constexpr int per    = TPt::integ_scratch;
  static constexpr int smem_doubles = 4 * MAX_Q * MAX_Q + EPB * per;
  static constexpr DerivativeLayout layout() { return make_layout(ct, ck, dim, d00, d01, d10, d11, (MASK & SB_MASK_SYM) != 0); }
};

#ifndef SB_EG_MINB
#define SB_EG_MINB 1
#endif
template <int dim, int q, int nk, int ck, int nt, int ct, int MASK, int EPB_ = 0>
__global__ void __launch_bounds__(ElemGradConfig<dim, q, nk, ck, nt, ct, MASK, EPB_>::threads, SB_EG_MINB)
    element_gradient_kernel(const __grid_constant__ ElemGradParams P)
{
  using C                       = ElemGradConfig<dim, q, nk, ck, nt, ct, MASK, EPB_>;
  constexpr DerivativeLayout LY = C::layout();
  constexpr int QPE = C::QPE, EPB = C::EPB, per = C::per;
  extern __shared__ double smem[];
  double* tables  = smem;
  double* scratch = smem + 4 * MAX_Q * MAX_Q;
  for (int i = threadIdx.x; i < 4 * MAX_Q * MAX_Q; i += C::threads) {
    const int           which = i / (2 * MAX_Q * MAX_Q);
    const int           rem   = i - which * (2 * MAX_Q * MAX_Q);
    const BasisTable1D& bt    = (which == 0) ? P.trial_basis : P.test_basis;
    tables[i]                 = (rem < MAX_Q * MAX_Q) ? bt.B[rem] : bt.G[rem - MAX_Q * MAX_Q];
  }
  __syncthreads();
  const double* Bk = tables;
  const double* Gk = tables + MAX_Q * MAX_Q;
  const double* Bt = tables + 2 * MAX_Q * MAX_Q;
  const double* Gt = tables + 3 * MAX_Q * MAX_Q;

  const int el = threadIdx.x / QPE;
  const int ql = threadIdx.x - el * QPE;
  const int qx = ql % q;
  const int qy = (ql / q) % q;
  const int qz = (dim == 3) ? ql / (q * q) : 0;

  const int  e_raw = blockIdx.x * EPB + el;
  const bool valid = e_raw < P.num_elements;
  const int  e     = valid ? e_raw : P.num_elements - 1;
  const int  elem  = P.elements ? P.elements[e] : e;

  const BasisTable1D& tb = P.test_basis;
  double              w  = tb.w[qx] * tb.w[qy];
  if (dim == 3) w *= tb.w[qz];

  // this quadrature point's derivative blocks, pre-multiplied by the quadrature weight
  const double*   dq     = (C::DQ_TE == 1) ? P.dqf + ((long long)blockIdx.x * EPB + el) * LY.entries * QPE + ql
                                           : P.dqf + (long long)blockIdx.x * LY.entries * C::threads + threadIdx.x;
  const long long stride = C::DQ_STRIDE;
  double          D[LY.entries > 0 ? LY.entries : 1];
  SB_UNROLL
  for (int i = 0; i < LY.entries; i++) D[i] = dq[i * stride] * w;

  constexpr bool has_source = C::d00 || C::d01;
  constexpr bool has_flux   = C::d10 || C::d11;
  constexpr int  cc         = ct * ck;
  constexpr int  ndt        = C::TPt::ndof;
  constexpr int  ndk        = C::ndof_k;

  for (int J = 0; J < ndk; J++) {
    const int jx = J % nk;
    const int jy = (J / nk) % nk;
    const int jz = (dim == 3) ? J / (nk * nk) : 0;
    double    phi, dphi[dim];
    if constexpr (dim == 3) {
      const double bx = Bk[qx * nk + jx], by = Bk[qy * nk + jy], bz = Bk[qz * nk + jz];
      phi     = bx * by * bz;
      dphi[0] = Gk[qx * nk + jx] * by * bz;
      dphi[1] = bx * Gk[qy * nk + jy] * bz;
      dphi[2] = bx * by * Gk[qz * nk + jz];
    } else {
      const double bx = Bk[qx * nk + jx], by = Bk[qy * nk + jy];
      phi     = bx * by;
      dphi[0] = Gk[qx * nk + jx] * by;
      dphi[1] = bx * Gk[qy * nk + jy];
    }
    // "component" index of the batched integrate: a * ck + b
    double src[cc];
    double flux[cc][dim];
    SB_UNROLL
    for (int a = 0; a < ct; a++) {
      SB_UNROLL
      for (int b = 0; b < ck; b++) {
        double s = 0.0;
        if constexpr (C::d00) s += D[LY.offset[0] + a * ck + b] * phi;
        if constexpr (C::d01) {
          SB_UNROLL
          for (int l = 0; l < dim; l++) s += D[LY.offset[1] + (a * ck + b) * dim + l] * dphi[l];
        }
        src[a * ck + b] = s;
        SB_UNROLL
        for (int k = 0; k < dim; k++) {
          double fl = 0.0;
          if constexpr (C::d10) fl += D[LY.offset[2] + (a * dim + k) * ck + b] * phi;
          if constexpr (C::d11) {
            SB_UNROLL
            for (int l = 0; l < dim; l++) fl += D[LY.i11(a, k, b, l)] * dphi[l];
          }
          flux[a * ck + b][k] = fl;
        }
      }
    }
    __syncthreads();
    double     out[cc];
    const bool owner = C::TPt::template integrate<has_source, has_flux>(Bt, Gt, scratch + el * per, qx, qy, qz, src,
                                                                        flux, out);
    if (owner && valid) {
      const int I = (dim == 3) ? (qz * nt + qy) * nt + qx : qy * nt + qx;
      SB_UNROLL
      for (int a = 0; a < ct; a++) {
        SB_UNROLL
        for (int b = 0; b < ck; b++) {
          const long long at = k_index(P.node_major, elem, b, J, a, I, ck, ndk, ct, ndt);
          P.K_e[at]          = P.overwrite ? out[a * ck + b] : P.K_e[at] + out[a * ck + b];
        }
      }
    }
  }
}

}  // namespace b200
}  // namespace serac
