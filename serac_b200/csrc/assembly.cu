// serac_b200: sparsity pattern and assembly of the gradient (include/serac_b200.h, "Sparsity pattern + assembly").
//
// Reference: GradientAssemblyLookupTables (F/dof_numbering.hpp:253-380) hashes every (row, col) pair of every element
// matrix and sorts them; Gradient::assemble (F/functional.hpp:586-652) then materialises all element matrices and adds
// their entries through a hash lookup each.  Here
//   * the pattern is built at NODE level -- per test node the sorted union of the trial nodes of its elements -- and
//     expanded by components (the expansion is where the vdof ordering of either space enters), bit-identical to the
//     reference's sorted unique pairs;
//   * assembly plans turn the loop around so that CSR rows are the unit of work (assemble_rows.cuh): no element
//     matrices in HBM, no atomics, every row written once.  Plans are host-built index tables, as the reference's
//     lookup tables are;
//   * integrals without a fused kernel (2-D, simplex, boundary faces) go through element matrices and an ordered
//     per-nonzero gather (sb_csr_accumulate).
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/scan.h>
#include <thrust/sequence.h>
#include <thrust/sort.h>

#include <algorithm>
#include <cstring>
#include <numeric>
#include <cstdlib>
#include <thread>

#include "internal.hpp"
#include "serac_b200/assemble_rows.cuh"

using namespace sb;
using serac::b200::AsmGroup;
using serac::b200::AsmNode;
using serac::b200::AsmUnit;

// ===================================================================================================================
// host side: node-level pattern, its expansion, assembly plans
// ===================================================================================================================
namespace {

template <typename F>
void parallel_for(int64_t n, int64_t grain, F&& f)
{
  const int64_t hw = std::max<unsigned>(1u, std::thread::hardware_concurrency());
  const int64_t nt = std::min<int64_t>(hw, (n + grain - 1) / std::max<int64_t>(grain, 1));
  if (nt <= 1) {
    f(int64_t(0), n);
    return;
  }
  std::vector<std::thread> th;
  const int64_t            chunk = (n + nt - 1) / nt;
  for (int64_t t = 0; t < nt; t++) {
    const int64_t b = t * chunk, e = std::min(n, b + chunk);
    if (b < e) th.emplace_back([=, &f] { f(b, e); });
  }
  for (auto& t : th) t.join();
}

/// node -> ascending list of (position in `elems`, local node): the transpose of the restriction table on a domain
struct Incidence {
  std::vector<int64_t> off;   // [num_nodes + 1]
  std::vector<int>     pos;   // position p of the element in the domain's list
  std::vector<int>     loc;   // local node index in that element
};

Incidence build_incidence(const int* table, int npe, int64_t num_nodes, const int* elems, int64_t nd)
{
  Incidence inc;
  inc.off.assign(size_t(num_nodes) + 1, 0);
  for (int64_t p = 0; p < nd; p++) {
    const int64_t e = elems ? elems[p] : p;
    for (int k = 0; k < npe; k++) inc.off[size_t(table[e * npe + k]) + 1]++;
  }
  for (int64_t i = 0; i < num_nodes; i++) inc.off[size_t(i) + 1] += inc.off[size_t(i)];
  inc.pos.resize(size_t(inc.off[size_t(num_nodes)]));
  inc.loc.resize(inc.pos.size());
  std::vector<int64_t> cur(inc.off.begin(), inc.off.end() - 1);
  for (int64_t p = 0; p < nd; p++) {  // p ascending: every node's list comes out sorted by p
    const int64_t e = elems ? elems[p] : p;
    for (int k = 0; k < npe; k++) {
      const int64_t at   = cur[size_t(table[e * npe + k])]++;
      inc.pos[size_t(at)] = int(p);
      inc.loc[size_t(at)] = k;
    }
  }
  return inc;
}

/// one block of a pair of BlockElementRestrictions: the tables of the two spaces on the elements of one geometry
struct PatternBlock {
  const int* test_table;
  int        npe_t;
  const int* trial_table;
  int        npe_k;
  int64_t    ne;
};

/// F/dof_numbering.hpp:290-351 at node level: row n = sorted unique trial nodes of the elements that contain test node n
/// (over every block: meshes of mixed element geometry have one block per geometry)
void build_node_pattern(const std::vector<PatternBlock>& blocks, int64_t num_test_nodes, std::vector<int64_t>& row_ptr,
                        std::vector<int>& cols)
{
  std::vector<Incidence> incs;
  for (const PatternBlock& b : blocks) incs.push_back(build_incidence(b.test_table, b.npe_t, num_test_nodes, nullptr, b.ne));
  row_ptr.assign(size_t(num_test_nodes) + 1, 0);
  // pass 1: row lengths, pass 2: fill (each row sorted + uniqued in a thread-local buffer)
  auto row_of = [&](int64_t n, std::vector<int>& buf) {
    buf.clear();
    for (size_t b = 0; b < blocks.size(); b++) {
      const Incidence&    inc = incs[b];
      const PatternBlock& B   = blocks[b];
      for (int64_t s = inc.off[size_t(n)]; s < inc.off[size_t(n) + 1]; s++) {
        const int64_t e = inc.pos[size_t(s)];
        if (s > inc.off[size_t(n)] && inc.pos[size_t(s) - 1] == e) continue;  // node repeated inside one element
        buf.insert(buf.end(), B.trial_table + e * B.npe_k, B.trial_table + (e + 1) * B.npe_k);
      }
    }
    std::sort(buf.begin(), buf.end());
    buf.erase(std::unique(buf.begin(), buf.end()), buf.end());
  };
  parallel_for(num_test_nodes, 4096, [&](int64_t b, int64_t e) {
    std::vector<int> buf;
    for (int64_t n = b; n < e; n++) {
      row_of(n, buf);
      row_ptr[size_t(n) + 1] = int64_t(buf.size());
    }
  });
  for (int64_t i = 0; i < num_test_nodes; i++) row_ptr[size_t(i) + 1] += row_ptr[size_t(i)];
  cols.resize(size_t(row_ptr[size_t(num_test_nodes)]));
  parallel_for(num_test_nodes, 4096, [&](int64_t b, int64_t e) {
    std::vector<int> buf;
    for (int64_t n = b; n < e; n++) {
      row_of(n, buf);
      std::copy(buf.begin(), buf.end(), cols.begin() + row_ptr[size_t(n)]);
    }
  });
}

void build_node_pattern(const int* test_table, int npe_t, int64_t num_test_nodes, const int* trial_table, int npe_k,
                        int64_t ne, std::vector<int64_t>& row_ptr, std::vector<int>& cols)
{
  build_node_pattern({PatternBlock{test_table, npe_t, trial_table, npe_k, ne}}, num_test_nodes, row_ptr, cols);
}

/// expansion of the node pattern by components: where each vdof row starts and which column sits at position j of it
struct Expansion {
  int     ct, ck, test_byvdim, trial_byvdim;
  int64_t NT, NK, nnz_nodes;
  /// first entry of the row (test node n, component a); len = trial nodes in the node row, first = node_row_ptr[n]
  __host__ __device__ int64_t row_start(int64_t first, int64_t len, int a) const
  {
    return test_byvdim ? (first * ct + a * len) * ck : (a * nnz_nodes + first) * ck;
  }
  /// position inside the row of (the j-th trial node of the node row, component b)
  __host__ __device__ int64_t col_pos(int64_t j, int64_t len, int b) const { return trial_byvdim ? j * ck + b : b * len + j; }
  __host__ __device__ int64_t col_of(int64_t m, int b) const { return trial_byvdim ? m * ck + b : b * NK + m; }
  __host__ __device__ int64_t row_of(int64_t n, int a) const { return test_byvdim ? n * ct + a : a * NT + n; }
};

struct HostPlan {
  std::vector<AsmGroup>       groups;
  std::vector<AsmUnit>        units;
  std::vector<AsmNode>        nodes;
  std::vector<unsigned short> pos;
  int                         acc_doubles = 0;
  bool                        covers_all_rows = false;
  std::string                 error;
};

/// groups / units / row positions for assemble_rows_kernel: see assemble_rows.cuh.  elems: the integral's domain
/// (NULL = all elements), nd entries; acc_cap: doubles of shared-memory accumulator a group may use
HostPlan build_plan(const int* test_table, int npe_t, int ct, int64_t NT, const int* trial_table, int npe_k, int ck,
                    const std::vector<int64_t>& nrp, const std::vector<int>& ncols, const Expansion& X,
                    const int* elems, int64_t nd, int acc_cap, int trial_blocks)
{
  HostPlan H;
  if (npe_t > 255) {
    H.error = "more than 255 nodes per element";
    return H;
  }
  trial_blocks = std::max(1, std::min(trial_blocks, 255));
  constexpr int item_cap = 96;  // AssembleRowsConfig::UCAP
  const Incidence inc = build_incidence(test_table, npe_t, NT, elems, nd);
  auto rowlen = [&](int64_t n) { return nrp[size_t(n) + 1] - nrp[size_t(n)]; };
  // rows of the pattern that exist vs rows this domain reaches
  int64_t rows_nonempty = 0, rows_reached = 0;
  for (int64_t n = 0; n < NT; n++) {
    rows_nonempty += rowlen(n) > 0 ? 1 : 0;
    rows_reached += inc.off[size_t(n) + 1] > inc.off[size_t(n)] ? 1 : 0;
  }
  H.covers_all_rows = (rows_nonempty == rows_reached);

  // position of every element trial node in the row of every element test node
  H.pos.resize(size_t(nd) * size_t(npe_t) * size_t(npe_k));
  bool too_long = false;
  parallel_for(nd, 256, [&](int64_t b, int64_t e) {
    for (int64_t p = b; p < e; p++) {
      const int64_t el = elems ? elems[p] : p;
      for (int I = 0; I < npe_t; I++) {
        const int64_t n     = test_table[el * npe_t + I];
        const int*    first = ncols.data() + nrp[size_t(n)];
        const int*    last  = ncols.data() + nrp[size_t(n) + 1];
        if (last - first > 65535) too_long = true;
        for (int J = 0; J < npe_k; J++) {
          const int* it = std::lower_bound(first, last, trial_table[el * npe_k + J]);
          H.pos[(size_t(p) * size_t(npe_t) + size_t(I)) * size_t(npe_k) + size_t(J)] = (unsigned short)(it - first);
        }
      }
    }
  });
  if (too_long) {
    H.error = "a node row has more than 65535 entries";
    return H;
  }

  // groups: one test component of the nodes whose first (lowest-position) element is p, in the element's local order, cut
  // to fit acc_cap.  An item = (element, local test node, trial block); the group's items are walked element-major
  std::vector<char> assigned(size_t(NT), 0);
  std::vector<int>  chunk;
  auto              flush = [&]() {
    if (chunk.empty()) return;
    for (int a = 0; a < ct; a++) {
      AsmGroup g{};
      g.a          = a;
      g.node_begin = int(H.nodes.size());
      g.unit_begin = int(H.units.size());
      int acc      = 0;
      for (size_t s = 0; s < chunk.size(); s++) {
        const int64_t n = chunk[s];
        AsmNode       nd_{};
        nd_.node    = int(n);
        nd_.acc_off = acc;
        nd_.rowlen  = int(rowlen(n));
        nd_.row_start = int(X.row_start(nrp[size_t(n)], nd_.rowlen, a));
        acc += nd_.rowlen * ck;
        H.nodes.push_back(nd_);
        for (int64_t t = inc.off[size_t(n)]; t < inc.off[size_t(n) + 1]; t++)
          for (int bg = 0; bg < trial_blocks; bg++)
            H.units.push_back(AsmUnit{inc.pos[size_t(t)], inc.loc[size_t(t)] | (bg << 8) | int(s << 16), -1, -1});
      }
      // element-major: a lane keeps the derivatives of (element, block) in registers across the nodes that share them,
      // and every row still sees its elements in ascending order
      std::sort(H.units.begin() + g.unit_begin, H.units.end(), [](const AsmUnit& x, const AsmUnit& y) {
        if (x.p != y.p) return x.p < y.p;
        const int kx = (x.meta >> 8) & 0xFF, ky = (y.meta >> 8) & 0xFF;  // trial block
        return kx != ky ? kx < ky : ((x.meta >> 16) & 0xFF) < ((y.meta >> 16) & 0xFF);
      });
      // look-ahead per item: the (element, block) key that follows this item's key, and the element of the key after
      const size_t ub = size_t(g.unit_begin), ue = H.units.size();
      auto key_of = [&](size_t i) { return (int64_t(H.units[i].p) << 8) | ((H.units[i].meta >> 8) & 0xFF); };
      int     n1_p = -1, n1_bg = 0, n2_p = -1;
      for (size_t i = ue; i-- > ub;) {
        if (i + 1 < ue && key_of(i + 1) != key_of(i)) {
          n2_p  = n1_p;
          n1_p  = H.units[i + 1].p;
          n1_bg = (H.units[i + 1].meta >> 8) & 0xFF;
        }
        H.units[i].next_p  = n1_p;
        H.units[i].next2_p = n2_p;
        H.units[i].meta |= n1_bg << 24;
      }
      g.node_end    = int(H.nodes.size());
      g.unit_end    = int(H.units.size());
      g.acc_total   = acc;
      H.acc_doubles = std::max(H.acc_doubles, acc);
      H.groups.push_back(g);
    }
    chunk.clear();
  };
  for (int64_t p = 0; p < nd; p++) {
    const int64_t el  = elems ? elems[p] : p;
    int           acc = 0, items = 0;
    for (int I = 0; I < npe_t; I++) {
      const int64_t n = test_table[el * npe_t + I];
      if (assigned[size_t(n)]) continue;
      assigned[size_t(n)] = 1;  // p ascending: the first element that reaches a node owns it
      const int need = int(rowlen(n)) * ck;  // one test component per group
      if (need > acc_cap) {
        H.error = "the rows of one node exceed the shared-memory accumulator";
        return H;
      }
      const int nitems = int(inc.off[size_t(n) + 1] - inc.off[size_t(n)]) * trial_blocks;
      if (nitems > item_cap) {
        H.error = "a node has more items than the kernel stages in shared memory";
        return H;
      }
      if (acc + need > acc_cap || items + nitems > item_cap || int(chunk.size()) == serac::b200::ASM_MAX_GROUP_NODES) {
        flush();
        acc = items = 0;
      }
      chunk.push_back(int(n));
      acc += need;
      items += nitems;
    }
    flush();
  }
  return H;
}

}  // namespace

extern "C" int sb_csr_pattern_host(const int* test_table, int npe_t, int ct, int64_t NT, int test_ordering,
                                   const int* trial_table, int npe_k, int ck, int64_t NK, int trial_ordering,
                                   int64_t ne, int* row_ptr, int* col_ind, int64_t* nnz_out)
{
  SB_REQUIRE(test_table && trial_table && row_ptr && nnz_out, "sb_csr_pattern_host: NULL argument");
  std::vector<int64_t> nrp;
  std::vector<int>     ncols;
  build_node_pattern(test_table, npe_t, NT, trial_table, npe_k, ne, nrp, ncols);
  Expansion X{ct, ck, test_ordering == SB_ORDERING_BYVDIM, trial_ordering == SB_ORDERING_BYVDIM, NT, NK, nrp[size_t(NT)]};
  const int64_t nnz = X.nnz_nodes * ct * ck;
  SB_REQUIRE(nnz < (int64_t(1) << 31), "sb_csr_pattern_host: more than 2^31 nonzeros");
  *nnz_out = nnz;
  for (int64_t n = 0; n < NT; n++) {
    const int64_t first = nrp[size_t(n)], len = nrp[size_t(n) + 1] - first;
    for (int a = 0; a < ct; a++) {
      const int64_t start      = X.row_start(first, len, a);
      row_ptr[X.row_of(n, a)] = int(start);
      if (!col_ind) continue;
      for (int64_t j = 0; j < len; j++)
        for (int b = 0; b < ck; b++) col_ind[start + X.col_pos(j, len, b)] = int(X.col_of(ncols[size_t(first + j)], b));
    }
  }
  row_ptr[NT * ct] = int(nnz);
  return SB_OK;
}

/// test hook (no device needed): run the index tables of an assembly plan on element matrices given by the caller,
/// exactly as assemble_rows_kernel walks them -- groups, units in order, accumulator offsets, row write-out.  It checks
/// the tables, not the arithmetic: K_e_host is [domain position][trial vdof][test vdof] (F/functional.hpp:599-601).
extern "C" int sb_debug_assembly_plan_host(const int* test_table, int npe_t, int ct, int64_t NT, int test_ordering,
                                           const int* trial_table, int npe_k, int ck, int64_t NK, int trial_ordering,
                                           int64_t ne, const int* elements, int64_t nd, int acc_cap, int trial_blocks,
                                           const double* K_e_host, const int* row_ptr, double* values, int* num_groups)
{
  std::vector<int64_t> nrp;
  std::vector<int>     ncols;
  build_node_pattern(test_table, npe_t, NT, trial_table, npe_k, ne, nrp, ncols);
  Expansion X{ct, ck, test_ordering == SB_ORDERING_BYVDIM, trial_ordering == SB_ORDERING_BYVDIM, NT, NK, nrp[size_t(NT)]};
  HostPlan H = build_plan(test_table, npe_t, ct, NT, trial_table, npe_k, ck, nrp, ncols, X, elements, nd, acc_cap, trial_blocks);
  SB_REQUIRE(H.error.empty(), H.error);
  if (num_groups) *num_groups = int(H.groups.size());
  const int64_t nrow = int64_t(npe_t) * ct, ncol = int64_t(npe_k) * ck;
  std::vector<double> acc(size_t(std::max(H.acc_doubles, 1)));
  for (const AsmGroup& g : H.groups) {
    std::fill(acc.begin(), acc.begin() + g.acc_total, 0.0);
    const int cpw = (ck + trial_blocks - 1) / trial_blocks;  // trial components per block
    for (int u = g.unit_begin; u < g.unit_end; u++) {
      const AsmUnit& un = H.units[size_t(u)];
      const int      I = un.meta & 0xFF, bg = (un.meta >> 8) & 0xFF, slot = (un.meta >> 16) & 0xFF;
      const AsmNode& nd_ = H.nodes[size_t(g.node_begin + slot)];
      for (int b = bg * cpw; b < std::min(ck, (bg + 1) * cpw); b++)
        for (int J = 0; J < npe_k; J++) {
          const int pj = H.pos[(size_t(un.p) * size_t(npe_t) + size_t(I)) * size_t(npe_k) + size_t(J)];
          acc[size_t(nd_.acc_off + X.col_pos(pj, nd_.rowlen, b))] +=
              K_e_host[(int64_t(un.p) * ncol + (b * npe_k + J)) * nrow + g.a * npe_t + I];
        }
    }
    for (int s = g.node_begin; s < g.node_end; s++) {
      const AsmNode& nd_ = H.nodes[size_t(s)];
      const int      len = nd_.rowlen * ck;
      if (nd_.row_start != row_ptr[X.row_of(nd_.node, g.a)]) return SB_ERR_INVALID;  // the plan's row starts are the pattern's
      for (int j = 0; j < len; j++) values[nd_.row_start + j] = acc[size_t(nd_.acc_off + j)];
    }
  }
  return SB_OK;
}

// ===================================================================================================================
// device side
// ===================================================================================================================
__global__ void expand_pattern_kernel(const long long* __restrict__ nrp, const int* __restrict__ ncols, Expansion X,
                                      int* __restrict__ row_ptr, int* __restrict__ col_ind)
{
  // one thread per test node: its ct rows
  for (long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x; n < X.NT;
       n += (long long)gridDim.x * blockDim.x) {
    const long long first = nrp[n], len = nrp[n + 1] - first;
    for (int a = 0; a < X.ct; a++) {
      const long long start   = X.row_start(first, len, a);
      row_ptr[X.row_of(n, a)] = int(start);
      for (long long j = 0; j < len; j++) {
        const long long m = ncols[first + j];
        for (int b = 0; b < X.ck; b++) col_ind[start + X.col_pos(j, len, b)] = int(X.col_of(m, b));
      }
    }
    if (n == 0) row_ptr[X.NT * X.ct] = int(X.nnz_nodes * X.ct * X.ck);
  }
}

extern "C" int sb_csr_create_blocks(int num_blocks, const sb_restriction_t* tests, const sb_restriction_t* trials,
                                    sb_csr_t* out)
{
  SB_REQUIRE(num_blocks >= 1 && tests && trials && out, "sb_csr_create_blocks: bad arguments");
  sb_restriction_t test = tests[0], trial = trials[0];
  SB_REQUIRE(test && trial, "sb_csr_create: NULL argument");
  std::vector<PatternBlock> blocks;
  for (int b = 0; b < num_blocks; b++) {
    sb_restriction_t t = tests[b], k = trials[b];
    SB_REQUIRE(t && k && t->num_elements == k->num_elements, "sb_csr_create: element counts differ");
    SB_REQUIRE(t->num_nodes == test->num_nodes && t->components == test->components && t->ordering == test->ordering &&
                   k->num_nodes == trial->num_nodes && k->components == trial->components && k->ordering == trial->ordering,
               "sb_csr_create_blocks: the blocks must belong to the same pair of spaces");
    blocks.push_back(PatternBlock{t->table_host.data(), t->npe, k->table_host.data(), k->npe, t->num_elements});
  }
  auto A      = std::make_unique<sb_csr_s>();
  A->num_rows = sb_restriction_lsize(test);
  A->num_cols = sb_restriction_lsize(trial);
  SB_REQUIRE(A->num_rows < (int64_t(1) << 31) && A->num_cols < (int64_t(1) << 31), "sb_csr_create: > 2^31 dofs");
  A->ct = test->components, A->ck = trial->components;
  A->test_ordering = test->ordering, A->trial_ordering = trial->ordering;
  A->num_test_nodes = test->num_nodes, A->num_trial_nodes = trial->num_nodes;
  A->test = test, A->trial = trial;
  build_node_pattern(blocks, test->num_nodes, A->node_row_ptr, A->node_cols);
  Expansion X{A->ct, A->ck, test->ordering == SB_ORDERING_BYVDIM, trial->ordering == SB_ORDERING_BYVDIM,
              test->num_nodes, trial->num_nodes, A->node_row_ptr[size_t(test->num_nodes)]};
  A->nnz = X.nnz_nodes * A->ct * A->ck;
  SB_REQUIRE(A->nnz < (int64_t(1) << 31), "sb_csr_create: more than 2^31 nonzeros (mfem::SparseMatrix indexes with int)");
  if (int e = dev_alloc(&A->row_ptr_dev, size_t(A->num_rows + 1))) return e;
  SB_CHECK_CUDA(cudaMemset(A->row_ptr_dev, 0, size_t(A->num_rows + 1) * sizeof(int)));
  if (int e = dev_alloc(&A->col_ind_dev, size_t(A->nnz))) return e;
  if (test->num_nodes > 0) {
    long long* nrp_dev   = nullptr;
    int*       ncols_dev = nullptr;
    static_assert(sizeof(long long) == sizeof(int64_t), "");
    if (int e = dev_upload(&nrp_dev, (const long long*)A->node_row_ptr.data(), A->node_row_ptr.size())) return e;
    if (int e = dev_upload(&ncols_dev, A->node_cols.data(), A->node_cols.size())) return e;
    expand_pattern_kernel<<<grid_for(test->num_nodes, 128), 128>>>(nrp_dev, ncols_dev, X, A->row_ptr_dev, A->col_ind_dev);
    SB_CHECK_CUDA(cudaGetLastError());
    SB_CHECK_CUDA(cudaDeviceSynchronize());
    cudaFree(nrp_dev);
    cudaFree(ncols_dev);
  }
  *out = A.release();
  return SB_OK;
}

extern "C" int sb_csr_create(sb_restriction_t test, sb_restriction_t trial, sb_csr_t* out)
{
  SB_REQUIRE(test && trial && out, "sb_csr_create: NULL argument");
  return sb_csr_create_blocks(1, &test, &trial, out);
}

// ---- transposed pattern (adjoint solves and sensitivities: HypreParMatrix::Transpose / MultTranspose,
//      src/serac/physics/solid_mechanics.hpp:1362-1371, 1404-1424) -------------------------------------------------------
__global__ void transpose_keys_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col_ind, long long num_rows,
                                      unsigned long long* __restrict__ keys, int* __restrict__ col_count)
{
  // one warp per row: keys[k] = (column << 32) | row; column histogram (integer atomics: order-independent)
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / 32;
  const int       lane = threadIdx.x % 32;
  const long long nw   = (long long)gridDim.x * blockDim.x / 32;
  for (long long r = warp; r < num_rows; r += nw) {
    for (int k = row_ptr[r] + lane; k < row_ptr[r + 1]; k += 32) {
      const int c = col_ind[k];
      keys[k]     = ((unsigned long long)(unsigned)c << 32) | (unsigned long long)(unsigned)r;
      atomicAdd(&col_count[c + 1], 1);
    }
  }
}

__global__ void transpose_cols_kernel(const unsigned long long* __restrict__ keys, long long nnz, int* __restrict__ col_ind_T)
{
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nnz; k += (long long)gridDim.x * blockDim.x)
    col_ind_T[k] = int(keys[k] & 0xFFFFFFFFull);
}

__global__ void transpose_values_kernel(const int* __restrict__ perm, const double* __restrict__ values,
                                        double* __restrict__ values_T, long long nnz)
{
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nnz; k += (long long)gridDim.x * blockDim.x)
    values_T[k] = values[perm[k]];
}

extern "C" int sb_csr_transpose(sb_csr_t A, sb_csr_t* out)
{
  SB_REQUIRE(A && out, "sb_csr_transpose: NULL argument");
  auto T      = std::make_unique<sb_csr_s>();
  T->nnz      = A->nnz;
  T->num_rows = A->num_cols;
  T->num_cols = A->num_rows;
  T->ct = A->ck, T->ck = A->ct;
  T->test_ordering = A->trial_ordering, T->trial_ordering = A->test_ordering;
  T->num_test_nodes = A->num_trial_nodes, T->num_trial_nodes = A->num_test_nodes;
  T->test = A->trial, T->trial = A->test;
  if (int e = dev_alloc(&T->row_ptr_dev, size_t(T->num_rows + 1))) return e;
  SB_CHECK_CUDA(cudaMemset(T->row_ptr_dev, 0, size_t(T->num_rows + 1) * sizeof(int)));
  if (int e = dev_alloc(&T->col_ind_dev, size_t(std::max<int64_t>(T->nnz, 1)))) return e;
  if (int e = dev_alloc(&T->transpose_perm_dev, size_t(std::max<int64_t>(T->nnz, 1)))) return e;
  if (A->nnz > 0) {
    unsigned long long* keys = nullptr;
    if (int e = dev_alloc(&keys, size_t(A->nnz))) return e;
    transpose_keys_kernel<<<grid_for(A->num_rows * 32, 256), 256>>>(A->row_ptr_dev, A->col_ind_dev, A->num_rows, keys,
                                                                     T->row_ptr_dev);
    SB_CHECK_CUDA(cudaGetLastError());
    thrust::device_ptr<unsigned long long> k(keys);
    thrust::device_ptr<int>                v(T->transpose_perm_dev);
    thrust::sequence(thrust::device, v, v + A->nnz);
    thrust::sort_by_key(thrust::device, k, k + A->nnz, v);  // keys are unique: (column, row) ascending
    thrust::device_ptr<int> rp(T->row_ptr_dev);
    thrust::inclusive_scan(thrust::device, rp, rp + T->num_rows + 1, rp);
    transpose_cols_kernel<<<grid_for(A->nnz, 256), 256>>>(keys, A->nnz, T->col_ind_dev);
    SB_CHECK_CUDA(cudaGetLastError());
    SB_CHECK_CUDA(cudaDeviceSynchronize());
    cudaFree(keys);
  }
  *out = T.release();
  return SB_OK;
}

extern "C" int sb_csr_transpose_values(sb_csr_t At, const double* values_dev, double* values_T_dev, void* stream)
{
  SB_REQUIRE(At && At->transpose_perm_dev, "sb_csr_transpose_values: the pattern was not made by sb_csr_transpose");
  if (At->nnz == 0) return SB_OK;
  SB_REQUIRE(values_dev && values_T_dev && values_dev != values_T_dev, "sb_csr_transpose_values: bad value arrays");
  transpose_values_kernel<<<grid_for(At->nnz, 256), 256, 0, cudaStream_t(stream)>>>(At->transpose_perm_dev, values_dev,
                                                                                      values_T_dev, At->nnz);
  sb_count_launch();
  SB_CHECK_CUDA(cudaGetLastError());
  return SB_OK;
}

static void free_plan(AssemblyPlan& P)
{
  cudaFree(P.groups_dev);
  cudaFree(P.units_dev);
  cudaFree(P.nodes_dev);
  cudaFree(P.pos_dev);
}

extern "C" int sb_csr_destroy(sb_csr_t A)
{
  if (!A) return SB_OK;
  for (auto& kv : A->extra) {
    cudaFree(kv.second.offsets_dev);
    cudaFree(kv.second.perm_dev);
    cudaFree(kv.second.nz_dev);
  }
  for (auto& kv : A->plans) free_plan(kv.second);
  cudaFree(A->row_ptr_dev);
  cudaFree(A->col_ind_dev);
  cudaFree(A->transpose_perm_dev);
  delete A;
  return SB_OK;
}

extern "C" int64_t sb_csr_nnz(sb_csr_t A) { return A->nnz; }
extern "C" int64_t sb_csr_num_rows(sb_csr_t A) { return A->num_rows; }
extern "C" int64_t sb_csr_num_cols(sb_csr_t A) { return A->num_cols; }
extern "C" int     sb_csr_pattern_dev(sb_csr_t A, const int** rp, const int** ci)
{
  *rp = A->row_ptr_dev;
  *ci = A->col_ind_dev;
  return SB_OK;
}
extern "C" int sb_csr_pattern_to_host(sb_csr_t A, int* rp, int* ci)
{
  SB_CHECK_CUDA(cudaMemcpy(rp, A->row_ptr_dev, size_t(A->num_rows + 1) * sizeof(int), cudaMemcpyDeviceToHost));
  if (A->nnz) SB_CHECK_CUDA(cudaMemcpy(ci, A->col_ind_dev, size_t(A->nnz) * sizeof(int), cudaMemcpyDeviceToHost));
  return SB_OK;
}

// ---- fused assembly ------------------------------------------------------------------------------------------------
#ifndef SB_ASM_ACC_CAP
#define SB_ASM_ACC_CAP 1600  // doubles: one component of the 8 node rows a p = 2 hex owns (1536) -- ~16 KB per warp
#endif

/// values (+)= the rows of integral I differentiated w.r.t. its trial `local_wrt`; fails with SB_ERR_NOT_FOUND when the
/// integral has no fused kernel or the plan cannot be built (the caller then goes through element matrices)
/// *written_inout: whether `values` already holds the sum of earlier integrals; a first integral whose plan reaches every
/// row stores (no memset, no read), otherwise `values` is zeroed once and rows are added
int csr_assemble_rows(sb_csr_t A, sb_integral_t I, int local_wrt, sb_restriction_t test, sb_restriction_t trial,
                      double* values, bool* written_inout, void* stream)
{
  if (!I->table.assemble_rows[local_wrt]) return SB_ERR_NOT_FOUND;
  if (test != A->test || trial != A->trial) return SB_ERR_NOT_FOUND;
  if (I->num_elements == 0) return SB_OK;
  SB_REQUIRE(I->dqf[local_wrt], "sb_functional_assemble: no stored derivative (call evaluate with this wrt first)");
  auto key = std::make_pair((const void*)I, local_wrt);
  auto it  = A->plans.find(key);
  if (it == A->plans.end()) {
    const int fixed = I->table.assemble_rows_fixed_doubles[local_wrt];
    int64_t   longest = 0;  // a group holds at least one whole row
    for (size_t n = 0; n + 1 < A->node_row_ptr.size(); n++)
      longest = std::max(longest, (A->node_row_ptr[n + 1] - A->node_row_ptr[n]) * A->ck);
    int64_t want = SB_ASM_ACC_CAP;
    if (const char* env = std::getenv("SB_ASM_ACC_CAP")) want = std::max(64, std::atoi(env));  // tuning experiments
    const int cap = int(std::min<int64_t>(std::max<int64_t>(want, longest), (220 * 1024) / 8 - fixed));
    const Expansion X{A->ct, A->ck, test->ordering == SB_ORDERING_BYVDIM, trial->ordering == SB_ORDERING_BYVDIM,
                      test->num_nodes, trial->num_nodes, A->node_row_ptr[size_t(test->num_nodes)]};
    HostPlan  H = build_plan(test->table_host.data(), test->npe, A->ct, test->num_nodes, trial->table_host.data(),
                             trial->npe, A->ck, A->node_row_ptr, A->node_cols, X,
                             I->entire_domain ? nullptr : I->elements_host.data(), I->num_elements, cap,
                             I->table.assemble_rows_trial_blocks[local_wrt]);
    if (!H.error.empty()) return SB_ERR_NOT_FOUND;
    AssemblyPlan P;
    P.num_groups      = int(H.groups.size());
    P.acc_doubles     = H.acc_doubles;
    P.covers_all_rows = H.covers_all_rows;
    if (int e = dev_upload((AsmGroup**)&P.groups_dev, H.groups.data(), H.groups.size())) return e;
    if (int e = dev_upload((AsmUnit**)&P.units_dev, H.units.data(), H.units.size())) return e;
    if (int e = dev_upload((AsmNode**)&P.nodes_dev, H.nodes.data(), H.nodes.size())) return e;
    if (int e = dev_upload(&P.pos_dev, H.pos.data(), H.pos.size())) return e;
    it = A->plans.emplace(key, P).first;
  }
  const AssemblyPlan&   P = it->second;
  sb_assemble_rows_args a{};
  a.dqf_dev         = I->dqf[local_wrt];
  a.groups_dev      = P.groups_dev;
  a.units_dev       = P.units_dev;
  a.nodes_dev       = P.nodes_dev;
  a.pos_dev         = P.pos_dev;
  a.row_ptr_dev     = A->row_ptr_dev;
  a.values_dev      = values;
  a.row_node_stride = test->node_stride;
  a.row_comp_stride = test->comp_stride;
  a.col_byvdim      = trial->ordering == SB_ORDERING_BYVDIM;
  if (!*written_inout && !P.covers_all_rows && A->nnz) {
    SB_CHECK_CUDA(cudaMemsetAsync(values, 0, size_t(A->nnz) * sizeof(double), cudaStream_t(stream)));
    *written_inout = true;
  }
  a.accumulate      = *written_inout ? 1 : 0;
  *written_inout    = true;
  a.num_groups      = P.num_groups;
  a.acc_doubles     = P.acc_doubles;
  int err = I->table.assemble_rows[local_wrt](&a, stream);
  if (err) {
    set_error(std::string("assemble_rows kernel launch: ") + cudaGetErrorString(cudaError_t(err)));
    return SB_ERR_CUDA;
  }
  return SB_OK;
}

// ---- element matrices -> values: ordered gather per nonzero (any geometry) -------------------------------------------
__global__ void emit_pairs_kernel(const int* __restrict__ test_table, const int* __restrict__ trial_table,
                                  unsigned long long* __restrict__ keys, long long total, int npe_t, int C_t,
                                  long long ns_t, long long cs_t, int npe_k, int C_k, long long ns_k, long long cs_k)
{
  const int nrow = npe_t * C_t, ncol = npe_k * C_k;
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < total;
       m += (long long)gridDim.x * blockDim.x) {
    // K_e(e, i, j): i = trial vdof (c*npe + node), j = test vdof      (F/functional.hpp:620-640)
    const long long e   = m / ((long long)nrow * ncol);
    const int       rem = int(m - e * (long long)nrow * ncol);
    const int       i   = rem / nrow;
    const int       j   = rem - i * nrow;
    const int       ck = i / npe_k, nk = i - ck * npe_k;
    const int       ct = j / npe_t, nt = j - ct * npe_t;
    const unsigned long long col = (unsigned long long)(trial_table[e * npe_k + nk] * ns_k + ck * cs_k);
    const unsigned long long row = (unsigned long long)(test_table[e * npe_t + nt] * ns_t + ct * cs_t);
    keys[m] = (row << 32) | col;
  }
}

__global__ void mark_heads_kernel(const unsigned long long* __restrict__ keys, long long total,
                                  long long* __restrict__ head_flag)
{
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < total;
       m += (long long)gridDim.x * blockDim.x) {
    head_flag[m] = (m == 0 || keys[m] != keys[m - 1]) ? 1 : 0;
  }
}

__global__ void extra_heads_kernel(const unsigned long long* __restrict__ keys, const long long* __restrict__ rank,
                                   long long total, const int* __restrict__ row_ptr, const int* __restrict__ col_ind,
                                   long long* __restrict__ offsets, long long* __restrict__ nz_of,
                                   int* __restrict__ missing)
{
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < total;
       m += (long long)gridDim.x * blockDim.x) {
    if (m == 0 || keys[m] != keys[m - 1]) {
      const long long u   = rank[m] - 1;
      const int       row = int(keys[m] >> 32), col = int(keys[m] & 0xFFFFFFFFull);
      int             lo = row_ptr[row], hi = row_ptr[row + 1] - 1, found = -1;  // col_ind is sorted within a row
      while (lo <= hi) {
        const int mid = (lo + hi) >> 1, c = col_ind[mid];
        if (c == col) {
          found = mid;
          break;
        }
        if (c < col)
          lo = mid + 1;
        else
          hi = mid - 1;
      }
      if (found < 0) atomicAdd(missing, 1);
      offsets[u] = m;
      nz_of[u]   = found;
    }
  }
}

__global__ void csr_accumulate_kernel(const long long* __restrict__ offsets, const uint32_t* __restrict__ perm,
                                      const long long* __restrict__ nz_of, const double* __restrict__ K_e,
                                      double* __restrict__ values, long long num_unique)
{
  for (long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x; u < num_unique;
       u += (long long)gridDim.x * blockDim.x) {
    double acc = values[nz_of[u]];  // one thread per nonzero: ordered, no atomics
    for (long long s = offsets[u]; s < offsets[u + 1]; s++) acc += K_e[perm[s]];
    values[nz_of[u]] = acc;
  }
}

static int csr_extra_map(sb_csr_t A, sb_restriction_t test, sb_restriction_t trial, CsrExtraMap** out)
{
  auto key = std::make_pair(test, trial);
  auto it  = A->extra.find(key);
  if (it != A->extra.end()) {
    *out = &it->second;
    return SB_OK;
  }
  SB_REQUIRE(test->num_elements == trial->num_elements, "assembly from element matrices: element counts differ");
  CsrExtraMap M;
  const int     nrow  = test->npe * test->components;
  const int     ncol  = trial->npe * trial->components;
  const int64_t total = test->num_elements * nrow * ncol;
  SB_REQUIRE(total < (int64_t(1) << 32), "assembly from element matrices: more than 2^32 element-matrix entries");
  M.num_entries = total;
  if (total > 0) {
    unsigned long long* keys = nullptr;
    long long*          rank = nullptr;
    int*                missing = nullptr;
    if (int e = dev_alloc(&keys, size_t(total))) return e;
    if (int e = dev_alloc(&M.perm_dev, size_t(total))) return e;
    if (int e = dev_alloc(&rank, size_t(total))) return e;
    if (int e = dev_alloc(&missing, 1)) return e;
    SB_CHECK_CUDA(cudaMemset(missing, 0, sizeof(int)));
    emit_pairs_kernel<<<grid_for(total, 256), 256>>>(test->table_dev, trial->table_dev, keys, total, test->npe,
                                                     test->components, test->node_stride, test->comp_stride, trial->npe,
                                                     trial->components, trial->node_stride, trial->comp_stride);
    SB_CHECK_CUDA(cudaGetLastError());
    thrust::device_ptr<unsigned long long> k(keys);
    thrust::device_ptr<uint32_t>           v(M.perm_dev);
    thrust::sequence(thrust::device, v, v + total);
    thrust::stable_sort_by_key(thrust::device, k, k + total, v);  // stable: ascending K_e order inside a nonzero
    mark_heads_kernel<<<grid_for(total, 256), 256>>>(keys, total, rank);
    SB_CHECK_CUDA(cudaGetLastError());
    thrust::device_ptr<long long> rk(rank);
    thrust::inclusive_scan(thrust::device, rk, rk + total, rk);
    long long nu = 0;
    SB_CHECK_CUDA(cudaMemcpy(&nu, rank + total - 1, sizeof(long long), cudaMemcpyDeviceToHost));
    M.num_unique = nu;
    if (int e = dev_alloc(&M.offsets_dev, size_t(nu + 1))) return e;
    if (int e = dev_alloc(&M.nz_dev, size_t(nu))) return e;
    extra_heads_kernel<<<grid_for(total, 256), 256>>>(keys, rank, total, A->row_ptr_dev, A->col_ind_dev,
                                                      (long long*)M.offsets_dev, (long long*)M.nz_dev, missing);
    SB_CHECK_CUDA(cudaGetLastError());
    SB_CHECK_CUDA(cudaMemcpy(M.offsets_dev + nu, &total, sizeof(int64_t), cudaMemcpyHostToDevice));
    int miss = 0;
    SB_CHECK_CUDA(cudaMemcpy(&miss, missing, sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(keys);
    cudaFree(rank);
    cudaFree(missing);
    // the reference's nz_LUT lookup would throw here (dof_numbering.hpp:358): face dof pairs are a subset of the
    // pairs of the element the face belongs to
    SB_REQUIRE(miss == 0, "assembly from element matrices: a dof pair is not in the sparsity pattern");
  }
  auto ins = A->extra.emplace(key, M);
  *out     = &ins.first->second;
  return SB_OK;
}

/// values[nz] += the entries of K_e (element matrices on `test` x `trial`: the domain restrictions or boundary-face
/// restrictions of the same spaces) in ascending (element, i, j) order
extern "C" int sb_csr_accumulate(sb_csr_t A, sb_restriction_t test, sb_restriction_t trial, const double* K_e,
                                 double* values, void* stream)
{
  SB_REQUIRE(A && test && trial, "sb_csr_accumulate: NULL argument");
  CsrExtraMap* M = nullptr;
  if (int e = csr_extra_map(A, test, trial, &M)) return e;
  if (M->num_unique == 0) return SB_OK;
  csr_accumulate_kernel<<<grid_for(M->num_unique, 256), 256, 0, cudaStream_t(stream)>>>(
      (const long long*)M->offsets_dev, M->perm_dev, (const long long*)M->nz_dev, K_e, values, M->num_unique);
  sb_count_launch();
  SB_CHECK_CUDA(cudaGetLastError());
  return SB_OK;
}

extern "C" int sb_csr_assemble(sb_csr_t A, const double* K_e, double* values, void* stream)
{
  SB_REQUIRE(A && values, "sb_csr_assemble: NULL argument");
  if (A->nnz == 0) return SB_OK;
  SB_CHECK_CUDA(cudaMemsetAsync(values, 0, size_t(A->nnz) * sizeof(double), cudaStream_t(stream)));
  return sb_csr_accumulate(A, A->test, A->trial, K_e, values, stream);
}
