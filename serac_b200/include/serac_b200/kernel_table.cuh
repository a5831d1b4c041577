// serac_b200: type erasure of the templated kernels into the C ABI's sb_kernel_table.
//
// The reference erases (geometry, Q, spaces, q-function) into three std::function tables per Integral
// (src/serac/numerics/functional/integral.hpp:127-145, generate_kernels :184-222).  A C ABI cannot carry a template or
// a std::function, so the same erasure is done here with plain function pointers: make_kernel_table<Sig, QF, State>()
// instantiates, for every trial argument i, evaluation_with_AD_[i], jvp_[i] and element_gradient_[i]
// (integral.hpp:205-221) and returns their launchers.  Any translation unit compiled by nvcc can do this for its own
// functors; registry_*.cu does it for the named integrands shipped in the library.
#pragma once

#include <cstdlib>
#include <cstring>

#include "serac_b200.h"
#include "serac_b200/assemble_rows.cuh"
#include "serac_b200/domain_integral.cuh"
#include "serac_b200/jvp_stream.cuh"
#include "serac_b200/simplex_integral.cuh"

namespace serac {
namespace b200 {

extern "C" void sb_count_launch(void);  // launch counter lives in the library (capi.cu)

inline FieldArg to_field(const sb_field_arg& f)
{
  return FieldArg{f.L_dev, f.table_dev, (long long)f.node_stride, (long long)f.comp_stride};
}

template <typename K>
inline int ensure_smem(K kernel, size_t bytes)
{
  if (bytes > 48 * 1024) {
    cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
    if (err != cudaSuccess) return int(err);
  }
  return 0;
}

/// q-functions may expose run-time flags as compile-time variants (num_variants, variant<V>, variant_index(),
/// as_variant<V>(): see MaterialStressFunctor); the launcher then picks the kernel specialised for the flag's value
template <typename QF, typename = void>
struct qf_variants {
  static constexpr int n = 1;
};
template <typename QF>
struct qf_variants<QF, std::void_t<decltype(QF::num_variants)>> {
  static constexpr int n = QF::num_variants;
};

/// number of SMs of the current device (cached)
inline int sm_count()
{
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      n = 148;
  }
  return n;
}

/// SB_EVAL_PREFETCH=0 switches the L2 prefetch of later tiles off (for A/B measurements)
inline bool use_eval_prefetch()
{
  static const bool v = [] {
    const char* e = std::getenv("SB_EVAL_PREFETCH");
    return !(e && std::strcmp(e, "0") == 0);
  }();
  return v;
}

template <typename Sig, int WRT, typename QF, typename State, int ROWS, int MINB, bool FAST>
int launch_evaluation_kernel(const EvalParams<Sig::NT>& P_in, const QF& qf, int num_elements, void* stream)
{
  auto         kernel = &evaluation_kernel<Sig, WRT, QF, State, ROWS, MINB, FAST>;
  const size_t smem   = size_t(eval_smem_doubles<Sig>()) * sizeof(double);
  if (int e = ensure_smem(kernel, smem)) return e;
  using LAY      = EvalLayout<Sig>;
  const int grid = (num_elements + LAY::CE - 1) / LAY::CE;
  static int resident = 0;  // CTAs per SM of this instantiation
  if (resident == 0) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, LAY::threads, smem) != cudaSuccess) nb = 1;
    resident = nb > 0 ? nb : 1;
  }
  EvalParams<Sig::NT> P = P_in;
  P.prefetch_distance   = use_eval_prefetch() ? resident * sm_count() : 0;
  kernel<<<grid, LAY::threads, smem, cudaStream_t(stream)>>>(P, qf);
  sb_count_launch();
  return int(cudaGetLastError());
}

template <typename Sig, int WRT, typename QF, typename State, int ROWS = 1, int MINB = 1, bool FAST = true>
int launch_evaluate(const sb_eval_args* a, const void* qf_blob, void* stream)
{
  constexpr int NT = Sig::NT;
  EvalParams<NT> P{};
  auto           fill = [&]<int... I>(std::integer_sequence<int, I...>)
  {
    ((P.trial[I] = to_field(a->trial[I]), P.trial_basis[I] = make_basis_table(Sig::template trial_space<I>::order, Sig::q)),
     ...);
  };
  fill(std::make_integer_sequence<int, NT>{});
  P.test_basis   = make_basis_table(Sig::test_space::order, Sig::q);
  if constexpr (Sig::simplex) {
    auto tabs = [&]<int... I>(std::integer_sequence<int, I...>)
    {
      ((P.simplex_tab[I] = simplex_table_dev(Sig::geometry, Sig::template trial_space<I>::order, Sig::q)), ...);
    };
    tabs(std::make_integer_sequence<int, NT>{});
    P.simplex_tab[NT] = simplex_table_dev(Sig::geometry, Sig::test_space::order, Sig::q);
    for (int i = 0; i <= NT; i++)
      if (!P.simplex_tab[i]) return int(cudaErrorMemoryAllocation);
  }
  P.X_q          = a->X_q_dev;
  P.J_q          = a->J_q_dev;
  P.elements     = a->elements_dev;
  P.num_elements  = a->num_elements;
  P.first_element = a->first_element;
  P.r_E          = a->r_E_dev;
  P.node_major   = a->r_E_node_major;
  P.dqf          = a->dqf_dev;
  P.qdata        = a->qdata_dev;
  P.update_state = a->update_state;
  P.t            = a->t;
  if (a->num_elements - a->first_element <= 0) return 0;
  alignas(16) unsigned char qf_storage[sizeof(QF)];
  std::memcpy(qf_storage, qf_blob, sizeof(QF));
  const QF& qf = *reinterpret_cast<const QF*>(qf_storage);
  if constexpr (qf_variants<QF>::n > 1) {
    int  err      = int(cudaErrorInvalidValue);
    auto dispatch = [&]<int... V>(std::integer_sequence<int, V...>)
    {
      (((qf.variant_index() == V)
            ? void(err = launch_evaluation_kernel<Sig, WRT, typename QF::template variant<V>, State, ROWS, MINB, FAST>(
                       P, qf.template as_variant<V>(), a->num_elements - a->first_element, stream))
            : void()),
       ...);
    };
    dispatch(std::make_integer_sequence<int, qf_variants<QF>::n>{});
    return err;
  } else {
    return launch_evaluation_kernel<Sig, WRT, QF, State, ROWS, MINB, FAST>(P, qf, a->num_elements - a->first_element, stream);
  }
}

/// resident CTAs per SM asked of ptxas for the streaming Jacobian action
/// (symmetric tangent storage: the tile is 45 instead of 81 entries per point, shared memory admits more CTAs, and
/// the kernel is bound by the shared-memory pipe and latency rather than by HBM; tools/tune.py on config A: 2 -> 87 us,
/// 5 / 6 -> 80 us, 8 -> 91 us, profiles/r4b_tune_jvp.txt)
constexpr int jvp_min_blocks(int dim, int q, int mask)
{
  if (dim == 3 && (mask & SB_MASK_SYM) != 0) return q == 3 ? 5 : (q == 4 ? 3 : SB_JVP_MINB);
  return SB_JVP_MINB;
}

/// SB_JVP_KERNEL=legacy selects the one-tile-per-CTA kernel (for A/B measurements); default is the TMA-fed stream kernel
inline bool use_stream_jvp()
{
  static const bool v = [] {
    const char* e = std::getenv("SB_JVP_KERNEL");
    return !(e && std::strcmp(e, "legacy") == 0);
  }();
  return v;
}

template <typename Sig, int WRT, int MASK>
int launch_jvp(const sb_jvp_args* a, void* stream)
{
  using trial       = typename Sig::template trial_space<WRT>;
  constexpr int nk  = trial::order + 1;
  constexpr int ck  = trial::components;
  using C           = JvpConfig<Sig::dim, Sig::q, nk, ck, Sig::test_n, Sig::test_c, MASK, Sig::EPB>;
  if constexpr (MASK == 0) {
    return 0;  // structurally zero derivative: dr_E block stays as the caller zeroed it
  } else {
    JvpParams P{};
    P.trial        = to_field(a->trial);
    P.trial_basis  = make_basis_table(trial::order, Sig::q);
    P.test_basis   = make_basis_table(Sig::test_space::order, Sig::q);
    P.elements     = a->elements_dev;
    P.num_elements = a->num_elements;
    P.dqf          = a->dqf_dev;
    P.r_E          = a->r_E_dev;
    P.node_major   = a->r_E_node_major;
    if (a->num_elements <= 0) return 0;
    if constexpr (Sig::simplex) {
      using SX            = SimplexConfig<Sig::geometry, Sig::q, trial::order, ck, Sig::test_space::order, Sig::test_c, MASK, Sig::EPB>;
      P.simplex_trial_tab = simplex_table_dev(Sig::geometry, trial::order, Sig::q);
      P.simplex_test_tab  = simplex_table_dev(Sig::geometry, Sig::test_space::order, Sig::q);
      if (!P.simplex_trial_tab || !P.simplex_test_tab) return int(cudaErrorMemoryAllocation);
      auto         kernel = &simplex_jvp_kernel<Sig::geometry, Sig::q, trial::order, ck, Sig::test_space::order, Sig::test_c, MASK, Sig::EPB>;
      const size_t smem   = size_t(SX::jvp_smem_doubles) * sizeof(double);
      if (int e = ensure_smem(kernel, smem)) return e;
      kernel<<<(a->num_elements + SX::EPB - 1) / SX::EPB, SX::threads, smem, cudaStream_t(stream)>>>(P);
      sb_count_launch();
      return int(cudaGetLastError());
    } else {
    const int tiles = (a->num_elements + C::EPB - 1) / C::EPB;
    using SC        = JvpStreamConfig<Sig::dim, Sig::q, nk, ck, Sig::test_n, Sig::test_c, MASK, Sig::EPB>;
    if constexpr (SC::supported && SC::smem_bytes <= 227 * 1024) {
      if (use_stream_jvp()) {
        // (registry_tune.cu: experiment switches >= 100 name the resident CTAs asked of ptxas)
        constexpr int JM = (Sig::tune >= 100) ? Sig::tune - 100 : jvp_min_blocks(Sig::dim, Sig::q, MASK);
        auto       kernel = &jvp_stream_kernel<Sig::dim, Sig::q, nk, ck, Sig::test_n, Sig::test_c, MASK, Sig::EPB, JM>;
        static int resident = 0;  // CTAs per SM for this instantiation
        if (resident == 0) {
          if (int e = ensure_smem(kernel, SC::smem_bytes)) return e;
          int nb = 0;
          if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, SC::threads, SC::smem_bytes) != cudaSuccess) nb = 1;
          resident = nb > 0 ? nb : 1;
        }
        const int grid = tiles < resident * sm_count() ? tiles : resident * sm_count();
        kernel<<<grid, SC::threads, SC::smem_bytes, cudaStream_t(stream)>>>(P, tiles);
        sb_count_launch();
        return int(cudaGetLastError());
      }
    }
    auto         kernel = &jvp_kernel<Sig::dim, Sig::q, nk, ck, Sig::test_n, Sig::test_c, MASK, Sig::EPB>;
    const size_t smem   = size_t(C::smem_doubles) * sizeof(double);
    if (int e = ensure_smem(kernel, smem)) return e;
    kernel<<<tiles, C::threads, smem, cudaStream_t(stream)>>>(P);
    sb_count_launch();
    return int(cudaGetLastError());
    }
  }
}

template <typename Sig, int WRT, int MASK>
int launch_element_gradient(const sb_elem_grad_args* a, void* stream)
{
  using trial      = typename Sig::template trial_space<WRT>;
  constexpr int nk = trial::order + 1;
  constexpr int ck = trial::components;
  using C          = ElemGradConfig<Sig::dim, Sig::q, nk, ck, Sig::test_n, Sig::test_c, MASK, Sig::EPB>;
  if constexpr (MASK == 0) {
    return 0;
  } else {
    ElemGradParams P{};
    P.trial_basis  = make_basis_table(trial::order, Sig::q);
    P.test_basis   = make_basis_table(Sig::test_space::order, Sig::q);
    P.elements     = a->elements_dev;
    P.num_elements = a->num_elements;
    P.dqf          = a->dqf_dev;
    P.K_e          = a->K_e_dev;
    P.node_major   = a->K_e_node_major;
    P.overwrite    = a->overwrite;
    if (a->num_elements <= 0) return 0;
    if constexpr (Sig::simplex) {
      using SX            = SimplexConfig<Sig::geometry, Sig::q, trial::order, ck, Sig::test_space::order, Sig::test_c, MASK, Sig::EPB>;
      P.simplex_trial_tab = simplex_table_dev(Sig::geometry, trial::order, Sig::q);
      P.simplex_test_tab  = simplex_table_dev(Sig::geometry, Sig::test_space::order, Sig::q);
      if (!P.simplex_trial_tab || !P.simplex_test_tab) return int(cudaErrorMemoryAllocation);
      auto kernel = &simplex_element_gradient_kernel<Sig::geometry, Sig::q, trial::order, ck, Sig::test_space::order, Sig::test_c, MASK, Sig::EPB>;
      const size_t smem = size_t(SX::grad_smem_doubles) * sizeof(double);
      if (int e = ensure_smem(kernel, smem)) return e;
      kernel<<<(a->num_elements + SX::EPB - 1) / SX::EPB, SX::threads, smem, cudaStream_t(stream)>>>(P);
      sb_count_launch();
      return int(cudaGetLastError());
    } else {
      auto         kernel = &element_gradient_kernel<Sig::dim, Sig::q, nk, ck, Sig::test_n, Sig::test_c, MASK, Sig::EPB>;
      const size_t smem   = size_t(C::smem_doubles) * sizeof(double);
      if (int e = ensure_smem(kernel, smem)) return e;
      const int grid = (a->num_elements + C::EPB - 1) / C::EPB;
      kernel<<<grid, C::threads, smem, cudaStream_t(stream)>>>(P);
      sb_count_launch();
      return int(cudaGetLastError());
    }
  }
}

/// the fused assembly exists for hexahedra (sum factorisation in three stages); everything else goes through K_e
template <typename Sig, int MASK>
constexpr bool has_assemble_rows()
{
  return MASK != 0 && Sig::dim == 3 && !Sig::simplex && !Sig::boundary;
}

template <typename Sig, int WRT, int MASK>
int launch_assemble_rows(const sb_assemble_rows_args* a, void* stream)
{
  if constexpr (!has_assemble_rows<Sig, MASK>()) {
    return int(cudaErrorNotSupported);
  } else {
    using trial      = typename Sig::template trial_space<WRT>;
    constexpr int nk = trial::order + 1;
    constexpr int ck = trial::components;
    using C = AssembleRowsConfig<Sig::q, nk, ck, Sig::test_n, Sig::test_c, MASK, Sig::EPB>;
    if (a->num_groups <= 0) return 0;
    AssembleRowsParams P{};
    P.trial_basis = make_basis_table(trial::order, Sig::q);
    P.test_basis  = make_basis_table(Sig::test_space::order, Sig::q);
    P.dqf         = a->dqf_dev;
    P.groups      = static_cast<const AsmGroup*>(a->groups_dev);
    P.units       = static_cast<const AsmUnit*>(a->units_dev);
    P.nodes       = static_cast<const AsmNode*>(a->nodes_dev);
    P.pos         = a->pos_dev;
    P.row_ptr     = a->row_ptr_dev;
    P.values      = a->values_dev;
    P.row_ns      = a->row_node_stride;
    P.row_cs      = a->row_comp_stride;
    P.col_byvdim  = a->col_byvdim;
    P.accumulate  = a->accumulate;
    P.num_groups  = a->num_groups;
    auto         kernel = &assemble_rows_kernel<Sig::q, nk, ck, Sig::test_n, Sig::test_c, MASK, Sig::EPB>;
    const size_t smem   = size_t(C::fixed_doubles + a->acc_doubles) * sizeof(double);
    if (int e = ensure_smem(kernel, smem)) return e;
    kernel<<<a->num_groups, C::threads, smem, cudaStream_t(stream)>>>(P);
    sb_count_launch();
    return int(cudaGetLastError());
  }
}

/// 128 registers per thread: 65536 / 128 = 512 threads per SM
template <typename Sig>
constexpr int default_min_blocks()
{
  constexpr int t32 = ((EvalLayout<Sig>::threads + 31) / 32) * 32;
  return (512 / t32) > 0 ? (512 / t32) : 1;
}

/// the entries of the table that depend on how the derivative w.r.t. trial WRT is stored (DT = its DerivativeTraits)
template <typename Sig, int WRT, typename DT>
void set_derivative_launchers(sb_kernel_table& T)
{
  constexpr DerivativeLayout L = DT::layout();
  constexpr int mask = (L.has[0] ? 1 : 0) | (L.has[1] ? 2 : 0) | (L.has[2] ? 4 : 0) | (L.has[3] ? 8 : 0) |
                       (L.sym ? SB_MASK_SYM : 0);
  constexpr int ck   = Sig::template trial_space<WRT>::components;
  T.jvp[WRT]                = &launch_jvp<Sig, WRT, mask>;
  T.element_gradient[WRT]   = &launch_element_gradient<Sig, WRT, mask>;
  if constexpr (has_assemble_rows<Sig, mask>()) {
    using trialW = typename Sig::template trial_space<WRT>;
    T.assemble_rows[WRT] = &launch_assemble_rows<Sig, WRT, mask>;
    using AC = AssembleRowsConfig<Sig::q, trialW::order + 1, ck, Sig::test_n, Sig::test_c, mask, Sig::EPB>;
    T.assemble_rows_fixed_doubles[WRT] = AC::fixed_doubles;
    T.assemble_rows_trial_blocks[WRT]  = AC::NB;
  }
  T.derivative_mask[WRT]         = mask;
  T.derivative_entries[WRT]      = L.entries;
  T.derivative_full_entries[WRT] = make_layout(L.ct, L.ck, L.dim, L.has[0], L.has[1], L.has[2], L.has[3]).entries;
}

template <typename Sig, typename QF, typename State, int WRT, int ROWS_REQ, int MINB, bool FAST>
void fill_wrt(sb_kernel_table& T)
{
  using DT                   = DerivativeTraits<Sig, WRT, QF, State, FAST>;
  constexpr int ck   = Sig::template trial_space<WRT>::components;
  constexpr int ROWS = (ROWS_REQ > 0 && ck % ROWS_REQ == 0) ? ROWS_REQ : 1;
  // resident CTAs per SM asked of ptxas: kernels whose q-function differentiates itself fit 128 registers / thread
  // (measured: 263 -> 178 us on config A); dual-number kernels already spill at 255 and are left alone
  constexpr int MB = (MINB > 0) ? MINB : (DT::tangent ? default_min_blocks<Sig>() : 1);
  T.evaluate_ad[WRT]        = &launch_evaluate<Sig, WRT, QF, State, ROWS, MB, FAST>;
  set_derivative_launchers<Sig, WRT, DT>(T);
}

/// sb_kernel_table::specialize: the launchers of the functor's own variant (launch_evaluate dispatches the evaluation
/// kernel on the same variant_index(), so what it stores is what these read)
template <typename Sig, typename QF, typename State, bool FAST>
void specialize_table(void* table, const void* qf_blob)
{
  if constexpr (qf_variants<QF>::n > 1) {
    sb_kernel_table& T = *static_cast<sb_kernel_table*>(table);
    alignas(16) unsigned char qf_storage[sizeof(QF)];
    std::memcpy(qf_storage, qf_blob, sizeof(QF));
    const QF& qf   = *reinterpret_cast<const QF*>(qf_storage);
    auto per_wrt = [&]<int... I>(std::integer_sequence<int, I...>)
    {
      auto per_variant = [&]<int W, int... V>(std::integral_constant<int, W>, std::integer_sequence<int, V...>)
      {
        (((qf.variant_index() == V)
              ? set_derivative_launchers<Sig, W, DerivativeTraits<Sig, W, typename QF::template variant<V>, State, FAST>>(T)
              : void()),
         ...);
      };
      (per_variant(std::integral_constant<int, I>{}, std::make_integer_sequence<int, qf_variants<QF>::n>{}), ...);
    };
    per_wrt(std::make_integer_sequence<int, Sig::NT>{});
  }
}

/// State = serac::Nothing: q-function takes no state argument; serac::Empty: takes an Empty& it never touches;
/// anything else: trivially copyable per-quadrature-point struct
/// ROWS_REQ: gradient rows per derivative pass (1 = least registers, ck = one pass); MINB: min resident CTAs per SM
/// FAST: let q-functions that provide a flux tangent use it (false = dual numbers everywhere, the reference's scheme)
/// MINB: __launch_bounds__ min resident CTAs per SM (0 = choose: default_min_blocks for plain / tangent kernels)
template <typename Sig, typename QF, typename State, int ROWS_REQ = 1, int MINB = 0, bool FAST = true>
sb_kernel_table make_kernel_table()
{
  static_assert(std::is_trivially_copyable_v<QF>, "q-function functors are copied by value into kernel parameters");
  sb_kernel_table T{};
  T.geom            = int(Sig::geometry);
  T.dim             = Sig::dim;
  T.spatial_dim     = Sig::sdim;
  T.q               = Sig::q;
  T.num_trials      = Sig::NT;
  T.test_order      = Sig::test_space::order;
  T.test_components = Sig::test_c;
  auto fill         = [&]<int... I>(std::integer_sequence<int, I...>)
  {
    ((T.trial_order[I] = Sig::template trial_space<I>::order,
      T.trial_components[I] = Sig::template trial_space<I>::components, fill_wrt<Sig, QF, State, I, ROWS_REQ, MINB, FAST>(T)),
     ...);
  };
  fill(std::make_integer_sequence<int, Sig::NT>{});
  T.qf_bytes           = sizeof(QF);
  T.state_bytes        = (std::is_same_v<State, Nothing> || std::is_empty_v<State>) ? 0 : sizeof(State);
  T.elements_per_block = Sig::EPB;
  T.derivative_tile_elements = Sig::DQ_TE;
  T.evaluate           = &launch_evaluate<Sig, -1, QF, State, 1, (MINB > 0 ? MINB : default_min_blocks<Sig>()), FAST>;
  T.specialize         = (qf_variants<QF>::n > 1) ? &specialize_table<Sig, QF, State, FAST> : nullptr;
  return T;
}

/// kernel table of a boundary integral (MakeBoundaryIntegral, integral.hpp:322-350): QF is the user's boundary
/// q-function `(t, tuple{x, dX/dxi}, tuple{value, d/dxi}...) -> source`; Sig is a BoundarySignature on the face geometry.
/// The functor blob handed to sb_integral_create is the QF itself (BoundaryIntegrand<QF> has the same layout).
template <typename Sig, typename QF, int ROWS_REQ = 1, int MINB = 0>
sb_kernel_table make_boundary_kernel_table()
{
  static_assert(Sig::boundary, "make_boundary_kernel_table needs a BoundarySignature");
  static_assert(sizeof(BoundaryIntegrand<QF>) == sizeof(QF), "adapter must not change the functor layout");
  return make_kernel_table<Sig, BoundaryIntegrand<QF>, Nothing, ROWS_REQ, MINB, false>();
}

/// kernel table of a quantity of interest: Sig's test space is serac::QOI, QF returns the scalar integrand.  The functor
/// blob handed to sb_integral_create is the QF itself.  State as for make_kernel_table.
template <typename Sig, typename QF, typename State = Nothing, int ROWS_REQ = 1, int MINB = 0>
sb_kernel_table make_qoi_kernel_table()
{
  static_assert(Sig::test_space::family == Family::QOI, "make_qoi_kernel_table needs a Signature whose test space is QOI");
  static_assert(sizeof(QoiIntegrand<QF>) == sizeof(QF), "adapter must not change the functor layout");
  return make_kernel_table<Sig, QoiIntegrand<QF>, State, ROWS_REQ, MINB, false>();
}

}  // namespace b200
}  // namespace serac
