// serac_b200: the opaque handle types of include/serac_b200.h, shared by the library's translation units (capi.cu:
// restrictions, integrals, Functional orchestration, halo exchange; assembly.cu: sparsity pattern + assembly).
#pragma once

#include <algorithm>
#include <cstdint>
#include <map>
#include <memory>
#include <utility>
#include <vector>

#include "common.hpp"

template <typename T>
static int dev_alloc(T** p, size_t n)
{
  *p = nullptr;
  if (n == 0) return SB_OK;
  SB_CHECK_CUDA(cudaMalloc((void**)p, n * sizeof(T)));
  return SB_OK;
}

template <typename T>
static int dev_upload(T** p, const T* host, size_t n)
{
  if (int e = dev_alloc(p, n)) return e;
  if (n) SB_CHECK_CUDA(cudaMemcpy(*p, host, n * sizeof(T), cudaMemcpyHostToDevice));
  return SB_OK;
}

static inline int grid_for(long long n, int block)
{
  long long g = (n + block - 1) / block;
  return int(std::min<long long>(std::max<long long>(g, 1), 148LL * 32));
}

// ElementRestriction (F/element_restriction.hpp:150-215)
struct sb_restriction_s {
  int64_t num_elements = 0, num_nodes = 0;
  int     npe = 0, components = 0, ordering = SB_ORDERING_BYVDIM;
  int64_t node_stride = 0, comp_stride = 0;
  int*    table_dev   = nullptr;  // [ne][npe] node ids
  // transpose: node -> ascending list of E-slots (e*npe + k)
  int64_t* t_offsets_dev = nullptr;  // [num_nodes + 1]
  int*     t_slots_dev   = nullptr;  // [ne*npe]  (slot = e*npe + k; fits int32 up to 2^31 slots)
  std::vector<int> table_host;
};

// Integral (F/integral.hpp:24-168)
struct sb_integral_s {
  sb_kernel_table            table;
  std::vector<unsigned char> qf;
  int64_t                    num_elements = 0;
  int*                       elements_dev = nullptr;
  std::vector<int>           elements_host;  // domain element ids (identity for EntireDomain)
  bool                       entire_domain = false;
  const double*              X_q = nullptr;
  const double*              J_q = nullptr;
  void*                      qdata = nullptr;
  double*                    dqf[SB_MAX_TRIALS] = {nullptr};
  int                        nq = 0;
};

// ConformingProlongationOperator state of one space on one rank (mfem/fem/pfespace.cpp:3599-3663)
struct sb_halo_s {
  int64_t lsize = 0, tsize = 0, nsend = 0, nrecv = 0;
  int*    ltdof_dev = nullptr;
  int*    send_idx_dev = nullptr;  // P: true dofs packed for sharers.  P^T: received contributions add into these
  int*    recv_idx_dev = nullptr;  // P: local dofs filled from owners. P^T: local dofs packed for owners
  // P^T adds must be ordered: true dof -> ascending list of recv-buffer slots (deterministic, pfespace.cpp:3912-3931)
  int64_t* add_offsets_dev = nullptr;
  int*     add_slots_dev   = nullptr;
  double*  send_buf = nullptr;
  double*  recv_buf = nullptr;
  double*  own_send = nullptr;  // allocations of this object (send_buf / recv_buf may point at caller memory)
  double*  own_recv = nullptr;
  // peer-memory transport (sb_halo_enable_peer)
  int                 nn = 0;
  std::vector<int64_t> send_off, recv_off;
  struct PeerState*   peer_dev = nullptr;
  bool                peer_enabled = false;
  // multi-process setup (comm.cu): the ranks of the neighbours, this rank's own ring and the mapped rings of the others
  std::vector<int>    neighbor_ranks;
  double*             ipc_ring = nullptr;              // cudaMalloc'd by sb_halo_enable_peer_ipc (freed with the object)
  std::vector<void*>  ipc_mapped;                      // cudaIpcOpenMemHandle results (closed with the object)
  // owned-first local numbering (ltdof[i] == i below tsize, -1 above): the T-vector is the prefix of the L-vector and
  // P / P^T only touch the shared entries ("in place")
  bool    owned_first  = false;
  int*    shared_t_dev = nullptr;  // true dofs that receive contributions in P^T
  int64_t nshared      = 0;
};


/// map of the element matrices of one pair of restrictions (domain elements or boundary faces) into a pattern
struct CsrExtraMap {
  int64_t   num_entries = 0, num_unique = 0;
  int64_t*  offsets_dev = nullptr;  // [num_unique + 1] into perm
  uint32_t* perm_dev    = nullptr;  // K_e flat positions grouped by nonzero, ascending inside a group
  int64_t*  nz_dev      = nullptr;  // [num_unique] index of the group's nonzero in the pattern
};

/// what assemble_rows_kernel walks for one (integral domain, test restriction, trial restriction): assembly.cu
struct AssemblyPlan {
  int      num_groups = 0, acc_doubles = 0;
  bool     covers_all_rows = false;  // every row of the pattern belongs to a group (no memset needed before a store pass)
  void*    groups_dev = nullptr;
  void*    units_dev  = nullptr;
  void*    nodes_dev  = nullptr;
  unsigned short* pos_dev = nullptr;
};

// GradientAssemblyLookupTables (F/dof_numbering.hpp:253-380)
struct sb_csr_s {
  int64_t nnz = 0, num_rows = 0, num_cols = 0;
  int*    row_ptr_dev = nullptr;  // [num_rows + 1]
  int*    col_ind_dev = nullptr;  // [nnz]
  // the pattern at node level (one entry per coupled (test node, trial node) pair): the vdof pattern above is its
  // expansion by components; kept on the host for the assembly plans
  int     ct = 0, ck = 0, test_ordering = SB_ORDERING_BYVDIM, trial_ordering = SB_ORDERING_BYVDIM;
  int64_t num_test_nodes = 0, num_trial_nodes = 0;
  std::vector<int64_t> node_row_ptr;  // [num_test_nodes + 1]
  std::vector<int>     node_cols;     // sorted trial nodes of each row
  sb_restriction_t     test = nullptr, trial = nullptr;  // the restrictions the pattern was built from (not owned)
  std::map<std::pair<sb_restriction_t, sb_restriction_t>, CsrExtraMap> extra;
  std::map<std::pair<const void*, int>, AssemblyPlan>                  plans;  // keyed by (integral, wrt)
  int* transpose_perm_dev = nullptr;  // patterns made by sb_csr_transpose: values_T[k] = values[transpose_perm[k]]
};
