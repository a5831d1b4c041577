// serac_b200: the two derivative consumers on simplex elements (triangle / tetrahedron H1, p <= 3), plus the device
// cache of the simplex shape-function tables.
//
//   action_of_gradient   src/serac/numerics/functional/domain_integral_kernels.hpp:277-304
//   element_gradient     src/serac/numerics/functional/domain_integral_kernels.hpp:327-361
//   element ops          src/serac/numerics/functional/detail/triangle_H1.inl:233-336, tetrahedron_H1.inl:333-429
//
// Simplex elements have no tensor structure: interpolation and integration are dense (ndof x nq) contractions with
// the {phi, dphi/dxi} table held in shared memory.  Thread mapping and the stored-derivative layout
// ([tile][entry][lane], one lane per quadrature point) are the same as for the tensor-product kernels, so the
// evaluation kernel (domain_integral.cuh) is shared and only branches on Sig::simplex.
#pragma once

#include "serac_b200/domain_integral.cuh"
#include "serac_b200/simplex_table.cuh"

namespace serac {
namespace b200 {

template <Geometry geom, int Q, int pk, int ck, int pt, int ct, int MASK, int EPB_ = 0>
struct SimplexConfig {
  static constexpr int  dim     = dimension_of(geom);
  static constexpr int  nq      = num_quadrature_points(geom, Q);
  static constexpr int  ndk     = num_nodes(geom, pk);
  static constexpr int  ndt     = num_nodes(geom, pt);
  static constexpr int  EPB     = (EPB_ > 0) ? EPB_ : elements_per_block(geom, Q);
  static constexpr int  threads = EPB * nq;
  static constexpr bool d00 = (MASK & 1) != 0, d01 = (MASK & 2) != 0, d10 = (MASK & 4) != 0, d11 = (MASK & 8) != 0;
  static constexpr int  tab_k   = nq * ndk * (1 + dim);
  static constexpr int  tab_t   = nq * ndt * (1 + dim);
  static constexpr int  tables  = tab_k + tab_t + nq;  // + weights
  /// per-element scratch: gathered du, later the integration staging
  static constexpr int jvp_per  = cmax(ck * ndk, ct * (1 + dim) * nq);
  static constexpr int grad_per = ct * ck * (1 + dim) * nq;
  static constexpr int jvp_smem_doubles  = tables + EPB * jvp_per;
  static constexpr int grad_smem_doubles = tables + EPB * grad_per;
  static constexpr DerivativeLayout layout() { return make_layout(ct, ck, dim, d00, d01, d10, d11, (MASK & SB_MASK_SYM) != 0); }
};

template <typename C>
__device__ __forceinline__ void simplex_load_tables(double* tables, const double* __restrict__ tab_k,
                                                    const double* __restrict__ tab_t)
{
  for (int i = threadIdx.x; i < C::tab_k; i += C::threads) tables[i] = tab_k[i];
  for (int i = threadIdx.x; i < C::tab_t + C::nq; i += C::threads) tables[C::tab_k + i] = tab_t[i];
}

template <Geometry geom, int Q, int pk, int ck, int pt, int ct, int MASK, int EPB_ = 0>
__global__ void __launch_bounds__(SimplexConfig<geom, Q, pk, ck, pt, ct, MASK, EPB_>::threads)
    simplex_jvp_kernel(const __grid_constant__ JvpParams P)
{
  using C                       = SimplexConfig<geom, Q, pk, ck, pt, ct, MASK, EPB_>;
  constexpr DerivativeLayout LY = C::layout();
  constexpr int dim = C::dim, nq = C::nq, EPB = C::EPB, per = C::jvp_per;
  extern __shared__ double smem[];
  double* tables  = smem;
  double* scratch = smem + C::tables;
  simplex_load_tables<C>(tables, P.simplex_trial_tab, P.simplex_test_tab);

  const int  el    = threadIdx.x / nq;
  const int  ql    = threadIdx.x - el * nq;
  const int  e_raw = blockIdx.x * EPB + el;
  const bool valid = e_raw < P.num_elements;
  const int  e     = valid ? e_raw : P.num_elements - 1;
  const int  elem  = P.elements ? P.elements[e] : e;

  const double* dq = P.dqf + (long long)blockIdx.x * LY.entries * C::threads + threadIdx.x;

  // gather du (element_restriction.cpp:443-454)
  const FieldArg& f = P.trial;
  for (int i = threadIdx.x; i < EPB * ck * C::ndk; i += C::threads) {
    const int le   = i / (ck * C::ndk);
    const int rem  = i - le * (ck * C::ndk);
    const int a    = rem / C::ndk;
    const int node = rem - a * C::ndk;
    int       ee   = blockIdx.x * EPB + le;
    if (ee >= P.num_elements) ee = P.num_elements - 1;
    const int       ge  = P.elements ? P.elements[ee] : ee;
    const long long idx = (long long)f.table[(long long)ge * C::ndk + node];
    scratch[le * per + a * C::ndk + node] = f.L[idx * f.node_stride + a * f.comp_stride];
  }
  __syncthreads();
  double v[ck];
  double g[ck][dim];
  SimplexOps<dim, C::ndk, nq, ck>::interpolate(tables, scratch + el * per, ql, v, g);
  const double w = tables[C::tab_k + C::tab_t + ql];

  constexpr bool  has_source = C::d00 || C::d01;
  constexpr bool  has_flux   = C::d10 || C::d11;
  double          src[ct];
  double          flux[ct][dim];
  const long long stride = C::threads;
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
  __syncthreads();  // every thread has finished reading the gathered dofs: the scratch becomes the integration staging
  double     out[ct];
  const bool owner = SimplexOps<dim, C::ndt, nq, ct>::template integrate<has_source, has_flux>(
      tables + C::tab_k, scratch + el * per, ql, src, flux, out);
  if (owner && valid) {
    SB_UNROLL
    for (int a = 0; a < ct; a++) P.r_E[e_index(P.node_major, elem, a, ql, ct, C::ndt)] = out[a];
  }
}

template <Geometry geom, int Q, int pk, int ck, int pt, int ct, int MASK, int EPB_ = 0>
__global__ void __launch_bounds__(SimplexConfig<geom, Q, pk, ck, pt, ct, MASK, EPB_>::threads)
    simplex_element_gradient_kernel(const __grid_constant__ ElemGradParams P)
{
  using C                       = SimplexConfig<geom, Q, pk, ck, pt, ct, MASK, EPB_>;
  constexpr DerivativeLayout LY = C::layout();
  constexpr int dim = C::dim, nq = C::nq, EPB = C::EPB, per = C::grad_per;
  extern __shared__ double smem[];
  double* tables  = smem;
  double* scratch = smem + C::tables;
  simplex_load_tables<C>(tables, P.simplex_trial_tab, P.simplex_test_tab);
  __syncthreads();

  const int  el    = threadIdx.x / nq;
  const int  ql    = threadIdx.x - el * nq;
  const int  e_raw = blockIdx.x * EPB + el;
  const bool valid = e_raw < P.num_elements;
  const int  e     = valid ? e_raw : P.num_elements - 1;
  const int  elem  = P.elements ? P.elements[e] : e;
  const double w   = tables[C::tab_k + C::tab_t + ql];

  const double*   dq     = P.dqf + (long long)blockIdx.x * LY.entries * C::threads + threadIdx.x;
  const long long stride = C::threads;
  double          D[LY.entries > 0 ? LY.entries : 1];
  SB_UNROLL
  for (int i = 0; i < LY.entries; i++) D[i] = dq[i * stride] * w;

  constexpr bool has_source = C::d00 || C::d01;
  constexpr bool has_flux   = C::d10 || C::d11;
  constexpr int  cc         = ct * ck;
  constexpr int  ndt = C::ndt, ndk = C::ndk;

  for (int J = 0; J < ndk; J++) {
    // trial shape function J at this thread's quadrature point (batch_apply_shape_fn, tetrahedron_H1.inl:333-357)
    const double* tj  = tables + (ql * ndk + J) * (1 + dim);
    const double  phi = tj[0];
    double        dphi[dim];
    SB_UNROLL
    for (int l = 0; l < dim; l++) dphi[l] = tj[1 + l];
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
    const bool owner = SimplexOps<dim, ndt, nq, cc>::template integrate<has_source, has_flux>(
        tables + C::tab_k, scratch + el * per, ql, src, flux, out);
    if (owner && valid) {
      SB_UNROLL
      for (int a = 0; a < ct; a++) {
        SB_UNROLL
        for (int b = 0; b < ck; b++) {
          const long long at = k_index(P.node_major, elem, b, J, a, ql, ck, ndk, ct, ndt);
          P.K_e[at]          = P.overwrite ? out[a * ck + b] : P.K_e[at] + out[a * ck + b];
        }
      }
    }
  }
}

}  // namespace b200
}  // namespace serac
