// serac_b200: the multi-rank part of the C ABI (include/serac_b200_comm.h): NCCL bootstrap, shared-dof exchange over
// ncclSend / ncclRecv, CUDA-IPC mapping of the neighbours' receive rings.  Host code only: the pack / unpack and
// peer-memory kernels are the ones in capi.cu.
//
// Reference counterparts: mfem::ConformingProlongationOperator::Mult / MultTranspose and its device variant
// (mfem/fem/pfespace.cpp:3599-3663, :3740-3933), applied by Functional around every evaluation
// (src/serac/numerics/functional/functional.hpp:452-454, :479).
#include <dlfcn.h>

#include <algorithm>
#include <cstring>

#include "internal.hpp"
#include "serac_b200_comm.h"

using namespace sb;

// ---- NCCL, loaded at run time --------------------------------------------------------------------------------------
namespace {
struct nccl_id {
  char internal[SB_COMM_ID_BYTES];
};
using nccl_comm = void*;
enum
{
  NCCL_CHAR   = 0,  // ncclInt8 / ncclChar
  NCCL_DOUBLE = 8,  // ncclFloat64
  NCCL_SUM    = 0
};

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(nccl_id*)                                                          = nullptr;
  int (*CommInitRank)(nccl_comm*, int, nccl_id, int)                                    = nullptr;
  int (*CommDestroy)(nccl_comm)                                                         = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, nccl_comm, cudaStream_t)            = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t)       = nullptr;
  int (*Send)(const void*, size_t, int, int, nccl_comm, cudaStream_t)                   = nullptr;
  int (*Recv)(void*, size_t, int, int, nccl_comm, cudaStream_t)                         = nullptr;
  int (*GroupStart)()                                                                   = nullptr;
  int (*GroupEnd)()                                                                     = nullptr;
  const char* (*GetErrorString)(int)                                                    = nullptr;
};

NcclApi* nccl()
{
  static NcclApi api;
  static bool    tried = false;
  if (tried) return api.lib ? &api : nullptr;
  tried             = true;
  const char* env   = std::getenv("SB_NCCL_LIB");
  const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n) continue;
    api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) return nullptr;
  auto sym = [&](const char* name) { return dlsym(api.lib, name); };
  api.GetUniqueId    = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank   = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.CommDestroy    = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.AllGather      = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
  api.AllReduce      = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
  api.Send           = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
  api.Recv           = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
  api.GroupStart     = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
  api.GroupEnd       = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.AllReduce || !api.Send ||
      !api.Recv || !api.GroupStart || !api.GroupEnd) {
    dlclose(api.lib);
    api.lib = nullptr;
    return nullptr;
  }
  return &api;
}

#define SB_NCCL(expr)                                                                                        \
  do {                                                                                                       \
    int _r = (expr);                                                                                         \
    if (_r != 0) {                                                                                           \
      NcclApi* _a = nccl();                                                                                  \
      set_error(std::string(#expr) + ": " + ((_a && _a->GetErrorString) ? _a->GetErrorString(_r) : "NCCL error")); \
      return SB_ERR_CUDA;                                                                                    \
    }                                                                                                        \
  } while (0)
}  // namespace

struct sb_comm_s {
  nccl_comm comm = nullptr;
  int       rank = 0, world = 1;
  double*   scratch = nullptr;  // one double on the device (barrier)
};

extern "C" int sb_comm_unique_id(void* id_bytes)
{
  SB_REQUIRE(id_bytes, "sb_comm_unique_id: NULL argument");
  NcclApi* N = nccl();
  SB_REQUIRE(N, "NCCL not found: libnccl.so.2 is not in the library search path (or set SB_NCCL_LIB)");
  nccl_id id;
  SB_NCCL(N->GetUniqueId(&id));
  std::memcpy(id_bytes, id.internal, SB_COMM_ID_BYTES);
  return SB_OK;
}

extern "C" int sb_comm_create(const void* id_bytes, int rank, int world, sb_comm_t* out)
{
  SB_REQUIRE(id_bytes && out && world >= 1 && rank >= 0 && rank < world, "sb_comm_create: bad arguments");
  NcclApi* N = nccl();
  SB_REQUIRE(N, "NCCL not found: libnccl.so.2 is not in the library search path (or set SB_NCCL_LIB)");
  auto    c = std::make_unique<sb_comm_s>();
  nccl_id id;
  std::memcpy(id.internal, id_bytes, SB_COMM_ID_BYTES);
  SB_NCCL(N->CommInitRank(&c->comm, world, id, rank));
  c->rank  = rank;
  c->world = world;
  SB_CHECK_CUDA(cudaMalloc((void**)&c->scratch, sizeof(double)));
  SB_CHECK_CUDA(cudaMemset(c->scratch, 0, sizeof(double)));
  *out = c.release();
  return SB_OK;
}

extern "C" int sb_comm_destroy(sb_comm_t c)
{
  if (!c) return SB_OK;
  if (c->comm && nccl()) nccl()->CommDestroy(c->comm);
  cudaFree(c->scratch);
  delete c;
  return SB_OK;
}

extern "C" int sb_comm_rank(sb_comm_t c) { return c->rank; }
extern "C" int sb_comm_size(sb_comm_t c) { return c->world; }

extern "C" int sb_comm_allreduce_sum(sb_comm_t c, double* values, int64_t n, void* stream)
{
  SB_REQUIRE(c && values && n >= 0, "sb_comm_allreduce_sum: bad arguments");
  if (n == 0) return SB_OK;
  SB_NCCL(nccl()->AllReduce(values, values, size_t(n), NCCL_DOUBLE, NCCL_SUM, c->comm, cudaStream_t(stream)));
  return SB_OK;
}

extern "C" int sb_comm_barrier(sb_comm_t c, void* stream)
{
  SB_REQUIRE(c, "sb_comm_barrier: NULL argument");
  SB_NCCL(nccl()->AllReduce(c->scratch, c->scratch, 1, NCCL_DOUBLE, NCCL_SUM, c->comm, cudaStream_t(stream)));
  SB_CHECK_CUDA(cudaStreamSynchronize(cudaStream_t(stream)));
  return SB_OK;
}

extern "C" int sb_halo_set_neighbor_ranks(sb_halo_t H, const int* ranks)
{
  SB_REQUIRE(H && (ranks || H->nn == 0), "sb_halo_set_neighbor_ranks: NULL argument");
  H->neighbor_ranks.assign(ranks, ranks + H->nn);
  return SB_OK;
}

// one round of point-to-point messages: segment n of the send buffer goes to neighbour n, segment n of the receive
// buffer comes from it; a single NCCL group, so the messages progress together
static int nccl_round(sb_halo_t H, sb_comm_t c, const std::vector<int64_t>& out_off, const std::vector<int64_t>& in_off,
                      cudaStream_t s)
{
  NcclApi* N = nccl();
  SB_REQUIRE(int(H->neighbor_ranks.size()) == H->nn, "shared-dof exchange over NCCL: call sb_halo_set_neighbor_ranks first");
  SB_NCCL(N->GroupStart());
  for (int n = 0; n < H->nn; n++) {
    const int64_t ns = out_off[size_t(n) + 1] - out_off[size_t(n)], nr = in_off[size_t(n) + 1] - in_off[size_t(n)];
    if (nr > 0) SB_NCCL(N->Recv(H->recv_buf + in_off[size_t(n)], size_t(nr), NCCL_DOUBLE, H->neighbor_ranks[size_t(n)], c->comm, s));
    if (ns > 0) SB_NCCL(N->Send(H->send_buf + out_off[size_t(n)], size_t(ns), NCCL_DOUBLE, H->neighbor_ranks[size_t(n)], c->comm, s));
  }
  SB_NCCL(N->GroupEnd());
  return SB_OK;
}

extern "C" int sb_halo_P_nccl(sb_halo_t H, sb_comm_t c, const double* T, double* L, void* stream)
{
  SB_REQUIRE(H && c && T && L, "sb_halo_P_nccl: NULL argument");
  if (int e = sb_halo_P_pack(H, T, stream)) return e;
  if (H->nn)
    if (int e = nccl_round(H, c, H->send_off, H->recv_off, cudaStream_t(stream))) return e;
  return sb_halo_P_finish(H, T, L, stream);
}

extern "C" int sb_halo_Pt_nccl(sb_halo_t H, sb_comm_t c, const double* L, double* T, void* stream)
{
  SB_REQUIRE(H && c && T && L, "sb_halo_Pt_nccl: NULL argument");
  if (int e = sb_halo_Pt_pack(H, L, stream)) return e;
  if (H->nn)
    if (int e = nccl_round(H, c, H->recv_off, H->send_off, cudaStream_t(stream))) return e;  // P^T sends what P received
  return sb_halo_Pt_finish(H, L, T, stream);
}

// ---- peer rings over CUDA IPC ------------------------------------------------------------------------------------------
namespace {
struct RingAd {  // what every rank tells the others about one halo object
  cudaIpcMemHandle_t handle;
  long long          buffer_doubles;
  int                nn;
  int                neighbors[SB_HALO_MAX_NEIGHBORS];
  long long          send_off[SB_HALO_MAX_NEIGHBORS + 1], recv_off[SB_HALO_MAX_NEIGHBORS + 1];
};
}  // namespace

extern "C" int sb_halo_enable_peer_ipc(sb_halo_t H, sb_comm_t c)
{
  SB_REQUIRE(H && c, "sb_halo_enable_peer_ipc: NULL argument");
  SB_REQUIRE(H->nn <= SB_HALO_MAX_NEIGHBORS, "sb_halo_enable_peer_ipc: too many neighbours");
  SB_REQUIRE(int(H->neighbor_ranks.size()) == H->nn, "sb_halo_enable_peer_ipc: call sb_halo_set_neighbor_ranks first");
  NcclApi* N = nccl();
  // 1. the slot size every ring uses = the largest message buffer of any rank
  double* slot_dev = nullptr;
  if (int e = dev_alloc(&slot_dev, size_t(c->world))) return e;
  SB_CHECK_CUDA(cudaMemset(slot_dev, 0, sizeof(double) * size_t(c->world)));
  const double mine = double(std::max<int64_t>(sb_halo_buffer_doubles(H), 1));
  SB_CHECK_CUDA(cudaMemcpy(slot_dev + c->rank, &mine, sizeof(double), cudaMemcpyHostToDevice));
  SB_NCCL(N->AllReduce(slot_dev, slot_dev, size_t(c->world), NCCL_DOUBLE, NCCL_SUM, c->comm, nullptr));
  std::vector<double> slots(size_t(c->world));
  SB_CHECK_CUDA(cudaMemcpy(slots.data(), slot_dev, sizeof(double) * slots.size(), cudaMemcpyDeviceToHost));
  cudaFree(slot_dev);
  const int64_t slot = int64_t(*std::max_element(slots.begin(), slots.end()));
  // 2. this rank's ring: two slots of (2 * slot + SB_HALO_MAX_NEIGHBORS) 8-byte words (flag and LL protocols), zeroed
  const size_t ring_doubles = size_t(4 * slot + 2 * SB_HALO_MAX_NEIGHBORS);
  cudaFree(H->ipc_ring);
  SB_CHECK_CUDA(cudaMalloc((void**)&H->ipc_ring, ring_doubles * sizeof(double)));
  SB_CHECK_CUDA(cudaMemset(H->ipc_ring, 0, ring_doubles * sizeof(double)));
  // 3. advertise handle + neighbour tables
  RingAd ad{};
  SB_CHECK_CUDA(cudaIpcGetMemHandle(&ad.handle, H->ipc_ring));
  ad.buffer_doubles = sb_halo_buffer_doubles(H);
  ad.nn             = H->nn;
  for (int n = 0; n < H->nn; n++) ad.neighbors[n] = H->neighbor_ranks[size_t(n)];
  for (int n = 0; n <= H->nn && H->nn > 0; n++) {
    ad.send_off[n] = H->send_off[size_t(n)];
    ad.recv_off[n] = H->recv_off[size_t(n)];
  }
  char* ads_dev = nullptr;
  if (int e = dev_alloc(&ads_dev, sizeof(RingAd) * size_t(c->world))) return e;
  SB_CHECK_CUDA(cudaMemcpy(ads_dev + sizeof(RingAd) * size_t(c->rank), &ad, sizeof(RingAd), cudaMemcpyHostToDevice));
  SB_NCCL(N->AllGather(ads_dev + sizeof(RingAd) * size_t(c->rank), ads_dev, sizeof(RingAd), NCCL_CHAR, c->comm, nullptr));
  std::vector<RingAd> ads(size_t(c->world));
  SB_CHECK_CUDA(cudaMemcpy(ads.data(), ads_dev, sizeof(RingAd) * ads.size(), cudaMemcpyDeviceToHost));
  cudaFree(ads_dev);
  // 4. map the neighbours' rings and find this rank's segments inside them
  for (void* p : H->ipc_mapped) cudaIpcCloseMemHandle(p);
  H->ipc_mapped.clear();
  double*   peer_ring[SB_HALO_MAX_NEIGHBORS] = {nullptr};
  int64_t   P_off[SB_HALO_MAX_NEIGHBORS] = {0}, Pt_off[SB_HALO_MAX_NEIGHBORS] = {0};
  int       my_idx[SB_HALO_MAX_NEIGHBORS] = {0};
  for (int n = 0; n < H->nn; n++) {
    const int     r  = H->neighbor_ranks[size_t(n)];
    const RingAd& ra = ads[size_t(r)];
    int           j  = -1;
    for (int k = 0; k < ra.nn; k++)
      if (ra.neighbors[k] == c->rank) j = k;
    SB_REQUIRE(j >= 0, "sb_halo_enable_peer_ipc: a neighbour does not list this rank as its neighbour");
    void* mapped = nullptr;
    SB_CHECK_CUDA(cudaIpcOpenMemHandle(&mapped, ra.handle, cudaIpcMemLazyEnablePeerAccess));
    H->ipc_mapped.push_back(mapped);
    peer_ring[n] = static_cast<double*>(mapped);
    P_off[n]     = ra.recv_off[j];  // where this rank's P segment starts in the neighbour's slot
    Pt_off[n]    = ra.send_off[j];
    my_idx[n]    = j;
  }
  if (int e = sb_halo_enable_peer(H, H->ipc_ring, slot, peer_ring, P_off, Pt_off, my_idx)) return e;
  // 5. nobody writes into a ring before every rank has zeroed and published its own
  return sb_comm_barrier(c, nullptr);
}


// ===================================================================================================================
// K = P^T A_L P in sub-assembled form and CG on it (include/serac_b200_comm.h)
// ===================================================================================================================
#include "serac_b200_solvers.h"

struct sb_par_csr_s {
  sb_comm_t     c = nullptr;
  sb_halo_t     H = nullptr;
  sb_csr_t      A = nullptr;
  const double* values = nullptr;
  double*       xL = nullptr;
  double*       yL = nullptr;
  double*       scalar = nullptr;  // device scalar of the all-reduced inner products
  double*       work[3] = {nullptr, nullptr, nullptr};  // r, z, d of the solver (T-sized)
};

namespace {
__global__ void zero_ghost_diagonal_kernel(const int* __restrict__ ess, long long num_ess, const int* __restrict__ ltdof,
                                           const int* __restrict__ row_ptr, const int* __restrict__ col_ind,
                                           double* __restrict__ values)
{
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < num_ess; i += (long long)gridDim.x * blockDim.x) {
    const int b = ess[i];
    if (ltdof[b] >= 0) continue;  // the owner's copy keeps the unit diagonal
    int lo = row_ptr[b], hi = row_ptr[b + 1] - 1;
    while (lo <= hi) {
      const int mid = (lo + hi) >> 1, cc = col_ind[mid];
      if (cc == b) {
        values[mid] = 0.0;
        break;
      }
      if (cc < b)
        lo = mid + 1;
      else
        hi = mid - 1;
    }
  }
}
__global__ void scale_kernel(double* __restrict__ z, const double* __restrict__ dinv, const double* __restrict__ r, long long n)
{
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    z[i] = dinv ? dinv[i] * r[i] : r[i];
}
__global__ void cg_step_kernel(double* __restrict__ x, double* __restrict__ r, const double* __restrict__ d,
                               const double* __restrict__ z, double alpha, long long n)
{
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    x[i] += alpha * d[i];
    r[i] -= alpha * z[i];
  }
}
__global__ void cg_direction_kernel2(double* __restrict__ d, const double* __restrict__ z, double beta, long long n)
{
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    d[i] = z[i] + beta * d[i];
}
}  // namespace

extern "C" int sb_par_csr_create(sb_comm_t c, sb_halo_t H, sb_csr_t A, double* values, const int* ess_host, int64_t num_ess,
                                 sb_par_csr_t* out)
{
  SB_REQUIRE(c && H && A && values && out && num_ess >= 0, "sb_par_csr_create: bad arguments");
  SB_REQUIRE(A->num_rows == H->lsize && A->num_cols == H->lsize, "sb_par_csr_create: the matrix must be L x L of the halo's space");
  auto K    = std::make_unique<sb_par_csr_s>();
  K->c      = c;
  K->H      = H;
  K->A      = A;
  K->values = values;
  if (num_ess > 0) {
    SB_REQUIRE(ess_host != nullptr, "sb_par_csr_create: NULL essential dof list");
    int* ess_dev = nullptr;
    if (int e = dev_upload(&ess_dev, ess_host, size_t(num_ess))) return e;
    if (int e = sb_csr_eliminate_rows_cols(A, values, ess_dev, num_ess, nullptr, nullptr)) return e;
    zero_ghost_diagonal_kernel<<<grid_for(num_ess, 128), 128>>>(ess_dev, num_ess, H->ltdof_dev, A->row_ptr_dev, A->col_ind_dev,
                                                                values);
    SB_CHECK_CUDA(cudaGetLastError());
    SB_CHECK_CUDA(cudaDeviceSynchronize());
    cudaFree(ess_dev);
  }
  if (int e = dev_alloc(&K->xL, size_t(std::max<int64_t>(H->lsize, 1)))) return e;
  if (int e = dev_alloc(&K->yL, size_t(std::max<int64_t>(H->lsize, 1)))) return e;
  if (int e = dev_alloc(&K->scalar, 1)) return e;
  for (auto& w : K->work)
    if (int e = dev_alloc(&w, size_t(std::max<int64_t>(H->tsize, 1)))) return e;
  *out = K.release();
  return SB_OK;
}

extern "C" int sb_par_csr_destroy(sb_par_csr_t K)
{
  if (!K) return SB_OK;
  cudaFree(K->xL);
  cudaFree(K->yL);
  cudaFree(K->scalar);
  for (auto w : K->work) cudaFree(w);
  delete K;
  return SB_OK;
}

extern "C" int sb_par_csr_mult(sb_par_csr_t K, const double* x_T, double* y_T, void* stream)
{
  SB_REQUIRE(K && x_T && y_T, "sb_par_csr_mult: NULL argument");
  if (int e = sb_halo_P_nccl(K->H, K->c, x_T, K->xL, stream)) return e;
  if (int e = sb_csr_spmv(K->A, K->values, K->xL, K->yL, stream)) return e;
  return sb_halo_Pt_nccl(K->H, K->c, K->yL, y_T, stream);
}

extern "C" int sb_par_csr_diagonal(sb_par_csr_t K, double* diag_T, void* stream)
{
  SB_REQUIRE(K && diag_T, "sb_par_csr_diagonal: NULL argument");
  if (int e = sb_csr_diagonal(K->A, K->values, K->yL, stream)) return e;
  return sb_halo_Pt_nccl(K->H, K->c, K->yL, diag_T, stream);
}

static int par_dot(sb_comm_t c, double* scalar_dev, const double* x, const double* y, int64_t n, double* result, void* stream)
{
  double local = 0.0;
  if (n > 0)
    if (int e = sb_vector_dot(x, y, n, &local, stream)) return e;
  cudaStream_t s = cudaStream_t(stream);
  SB_CHECK_CUDA(cudaMemcpyAsync(scalar_dev, &local, sizeof(double), cudaMemcpyHostToDevice, s));
  if (int e = sb_comm_allreduce_sum(c, scalar_dev, 1, stream)) return e;
  SB_CHECK_CUDA(cudaMemcpyAsync(result, scalar_dev, sizeof(double), cudaMemcpyDeviceToHost, s));
  SB_CHECK_CUDA(cudaStreamSynchronize(s));
  return SB_OK;
}

extern "C" int sb_par_dot(sb_comm_t c, const double* x_T, const double* y_T, int64_t tsize, double* result_host, void* stream)
{
  SB_REQUIRE(c && result_host && tsize >= 0, "sb_par_dot: bad arguments");
  double* scalar = nullptr;
  if (int e = dev_alloc(&scalar, 1)) return e;
  const int rc = par_dot(c, scalar, x_T, y_T, tsize, result_host, stream);
  cudaFree(scalar);
  return rc;
}

extern "C" int sb_par_cg_solve(sb_par_csr_t K, const double* dinv, const double* b, double* x, double rel_tol, double abs_tol,
                               int max_iter, int* converged, int* final_iter, double* final_norm, void* stream)
{
  SB_REQUIRE(K && b && x && converged && final_iter && final_norm, "sb_par_cg_solve: NULL argument");
  cudaStream_t  s = cudaStream_t(stream);
  const int64_t n = K->H->tsize;
  double *      r = K->work[0], *z = K->work[1], *d = K->work[2];
  const int     grid = grid_for(std::max<int64_t>(n, 1), 256);
  *converged = 0, *final_iter = 0, *final_norm = 0.0;
  // mfem::CGSolver::Mult with iterative_mode = false (solvers.cpp:737-748): x = 0, r = b, z = B r, d = z
  if (n) SB_CHECK_CUDA(cudaMemsetAsync(x, 0, size_t(n) * sizeof(double), s));
  if (n) SB_CHECK_CUDA(cudaMemcpyAsync(r, b, size_t(n) * sizeof(double), cudaMemcpyDeviceToDevice, s));
  scale_kernel<<<grid, 256, 0, s>>>(z, dinv, r, n);
  if (n) SB_CHECK_CUDA(cudaMemcpyAsync(d, z, size_t(n) * sizeof(double), cudaMemcpyDeviceToDevice, s));
  double nom = 0.0, den = 0.0, betanom = 0.0;
  if (int e = par_dot(K->c, K->scalar, d, r, n, &nom, stream)) return e;
  *final_norm = std::sqrt(std::max(nom, 0.0));
  if (nom < 0.0) return SB_OK;  // the preconditioner is not positive definite (:765-775)
  const double r0 = std::max(nom * rel_tol * rel_tol, abs_tol * abs_tol);
  if (nom <= r0) {
    *converged = 1;
    return SB_OK;
  }
  if (int e = sb_par_csr_mult(K, d, z, stream)) return e;
  if (int e = par_dot(K->c, K->scalar, z, d, n, &den, stream)) return e;
  if (den <= 0.0) return SB_OK;  // the operator is not positive definite (:790-802)
  for (int i = 1;; i++) {
    const double alpha = nom / den;
    cg_step_kernel<<<grid, 256, 0, s>>>(x, r, d, z, alpha, n);
    scale_kernel<<<grid, 256, 0, s>>>(z, dinv, r, n);
    if (int e = par_dot(K->c, K->scalar, r, z, n, &betanom, stream)) return e;
    *final_iter = i;
    *final_norm = std::sqrt(std::max(betanom, 0.0));
    if (betanom <= r0) {
      *converged = 1;
      return SB_OK;
    }
    if (i + 1 > max_iter) return SB_OK;
    const double beta = betanom / nom;
    cg_direction_kernel2<<<grid, 256, 0, s>>>(d, z, beta, n);
    if (int e = sb_par_csr_mult(K, d, z, stream)) return e;
    if (int e = par_dot(K->c, K->scalar, d, z, n, &den, stream)) return e;
    if (den <= 0.0) return SB_OK;
    nom = betanom;
  }
}
