// Two (or more) ranks of one mesh partition driven from C++ through the C ABI alone -- no Python, no torch, no MPI:
// NCCL bootstrap (sb_comm_unique_id / sb_comm_create), the shared-dof exchange P / P^T over ncclSend / ncclRecv
// (sb_halo_P_nccl / sb_halo_Pt_nccl) and over CUDA-IPC-mapped peer rings (sb_halo_enable_peer_ipc + the in-place LL
// kernels), around sb_functional_evaluate / sb_functional_jvp.  This is the orchestration the reference has in
// Functional::operator() (src/serac/numerics/functional/functional.hpp:446-498: P on every argument, kernels, P^T on the
// output) with mfem's ConformingProlongationOperator (mfem/fem/pfespace.cpp:3599-3663) underneath.
//
//   par_functional_test <directory>
// The directory is written by tests/test_gpu_cpp_multirank.py: meta.txt (world, sizes) and, per rank, the partition's
// restriction tables, vertex coordinates, halo lists and T-vectors; every rank writes its results back (rank<r>/out_*.bin)
// and pytest compares them with the same Functional evaluated on the whole mesh.  The process forks one child per rank
// before touching CUDA; rank 0 publishes the NCCL id through a file.
#include <cuda_runtime.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <string>
#include <vector>

#include "serac_b200.h"
#include "serac_b200_comm.h"
#include "serac_b200_solvers.h"

#define CHECK(expr)                                                                          \
  do {                                                                                       \
    int _rc = (expr);                                                                        \
    if (_rc != 0) {                                                                          \
      std::fprintf(stderr, "rank %d: %s failed (%d): %s\n", g_rank, #expr, _rc, sb_last_error()); \
      std::exit(10 + g_rank);                                                                \
    }                                                                                        \
  } while (0)
#define CUDA(expr)                                                                             \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      std::fprintf(stderr, "rank %d: %s: %s\n", g_rank, #expr, cudaGetErrorString(_e));        \
      std::exit(20 + g_rank);                                                                  \
    }                                                                                          \
  } while (0)

static int         g_rank = -1;
static std::string g_dir;

template <typename T>
std::vector<T> read_bin(const std::string& path)
{
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) {
    std::fprintf(stderr, "missing input %s\n", path.c_str());
    std::exit(2);
  }
  const std::streamsize n = f.tellg();
  f.seekg(0);
  std::vector<T> v(size_t(n) / sizeof(T));
  f.read(reinterpret_cast<char*>(v.data()), n);
  return v;
}
template <typename T>
void write_bin(const std::string& path, const std::vector<T>& v)
{
  std::ofstream f(path, std::ios::binary);
  f.write(reinterpret_cast<const char*>(v.data()), std::streamsize(v.size() * sizeof(T)));
}
std::map<std::string, double> read_meta(const std::string& path)
{
  std::map<std::string, double> m;
  std::ifstream                 f(path);
  std::string                   k;
  double                        v;
  while (f >> k >> v) m[k] = v;
  return m;
}
template <typename T>
T* to_device(const std::vector<T>& h)
{
  T* d = nullptr;
  CUDA(cudaMalloc(&d, sizeof(T) * (h.size() ? h.size() : 1)));
  if (!h.empty()) CUDA(cudaMemcpy(d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
  return d;
}
std::vector<double> to_host(const double* d, size_t n)
{
  std::vector<double> h(n);
  CUDA(cudaMemcpy(h.data(), d, sizeof(double) * n, cudaMemcpyDeviceToHost));
  return h;
}

struct HaloLists {
  std::vector<int>     ltdof, neighbors, send_idx, recv_idx;
  std::vector<int64_t> send_off, recv_off;
  int64_t              lsize, tsize;
};
HaloLists read_halo(const std::string& dir, const std::string& tag)
{
  HaloLists l;
  l.ltdof     = read_bin<int>(dir + "/" + tag + "_ltdof.bin");
  l.neighbors = read_bin<int>(dir + "/" + tag + "_neighbors.bin");
  l.send_idx  = read_bin<int>(dir + "/" + tag + "_send_idx.bin");
  l.recv_idx  = read_bin<int>(dir + "/" + tag + "_recv_idx.bin");
  l.send_off  = read_bin<int64_t>(dir + "/" + tag + "_send_off.bin");
  l.recv_off  = read_bin<int64_t>(dir + "/" + tag + "_recv_off.bin");
  l.lsize     = int64_t(l.ltdof.size());
  l.tsize     = 0;
  for (int t : l.ltdof) l.tsize = std::max<int64_t>(l.tsize, int64_t(t) + 1);
  return l;
}
sb_halo_t make_halo(const HaloLists& l)
{
  sb_halo_t H = nullptr;
  CHECK(sb_halo_create(l.lsize, l.tsize, l.ltdof.data(), int(l.neighbors.size()), l.send_off.data(), l.send_idx.data(),
                       l.recv_off.data(), l.recv_idx.data(), &H));
  CHECK(sb_halo_set_neighbor_ranks(H, l.neighbors.data()));
  return H;
}

int run_rank(int rank, int world)
{
  g_rank = rank;
  int ndev = 0;
  CUDA(cudaGetDeviceCount(&ndev));
  if (ndev < world) {
    std::fprintf(stderr, "rank %d: %d ranks need %d GPUs, %d visible\n", rank, world, world, ndev);
    return 3;
  }
  CUDA(cudaSetDevice(rank));
  // --- NCCL bootstrap: rank 0 publishes the id through a file ---
  const std::string id_path = g_dir + "/nccl_id.bin";
  std::vector<char> id(SB_COMM_ID_BYTES);
  if (rank == 0) {
    CHECK(sb_comm_unique_id(id.data()));
    write_bin(id_path + ".tmp", id);
    std::rename((id_path + ".tmp").c_str(), id_path.c_str());
  } else {
    for (int tries = 0; access(id_path.c_str(), R_OK) != 0; tries++) {
      if (tries > 6000) return 4;
      usleep(10000);
    }
    id = read_bin<char>(id_path);
  }
  sb_comm_t comm = nullptr;
  CHECK(sb_comm_create(id.data(), rank, world, &comm));

  const std::string rd   = g_dir + "/rank" + std::to_string(rank);
  auto              meta = read_meta(rd + "/meta.txt");
  const int64_t     ne = int64_t(meta["ne"]), nn = int64_t(meta["num_nodes"]), nn1 = int64_t(meta["num_nodes1"]);
  auto tab = read_bin<uint64_t>(rd + "/tab.bin"), tab1 = read_bin<uint64_t>(rd + "/tab1.bin");
  sb_restriction_t G = nullptr, G1 = nullptr;
  CHECK(sb_restriction_create(tab.data(), ne, 27, 3, nn, SB_ORDERING_BYVDIM, &G));
  CHECK(sb_restriction_create(tab1.data(), ne, 8, 3, nn1, SB_ORDERING_BYVDIM, &G1));
  double* coords = to_device(read_bin<double>(rd + "/coords.bin"));
  double *X_q = nullptr, *J_q = nullptr;
  CUDA(cudaMalloc(&X_q, sizeof(double) * size_t(ne) * 3 * 27));
  CUDA(cudaMalloc(&J_q, sizeof(double) * size_t(ne) * 9 * 27));
  CHECK(sb_geometric_factors(SB_GEOM_CUBE, 1, 3, G1, coords, nullptr, ne, X_q, J_q, nullptr));
  const sb_kernel_table* table = nullptr;
  CHECK(sb_named_kernel_table("solid_neohookean_hex_p2", &table));
  std::vector<unsigned char> blob(table->qf_bytes);
  const double params[4] = {1.0, 100.0, 50.0, 1.0};
  CHECK(sb_named_qf_blob("solid_neohookean_hex_p2", params, 4, blob.data(), blob.size()));
  sb_integral_t I = nullptr;
  CHECK(sb_integral_create(table, blob.data(), nullptr, ne, X_q, J_q, &I));
  sb_functional_t        F       = nullptr;
  const sb_restriction_t trials[3] = {G1, G, G};
  CHECK(sb_functional_create(G, 3, trials, &F));
  const int which[3] = {0, 1, 2};
  CHECK(sb_functional_add_integral(F, I, which));

  HaloLists l2 = read_halo(rd, "p2"), l1 = read_halo(rd, "p1");
  // T-vectors, stored as L-sized buffers whose prefix is the T-vector (owned-first numbering)
  auto hS = read_bin<double>(rd + "/S.bin"), hU = read_bin<double>(rd + "/U.bin"), hA = read_bin<double>(rd + "/A.bin"),
       hdU = read_bin<double>(rd + "/dU.bin");
  const size_t L2 = size_t(l2.lsize), T2 = size_t(l2.tsize);
  double *     r_L = nullptr, *r_T = nullptr;
  CUDA(cudaMalloc(&r_L, sizeof(double) * L2));
  CUDA(cudaMalloc(&r_T, sizeof(double) * T2));

  // ===== transport 1: NCCL send / recv =====
  {
    sb_halo_t Hs = make_halo(l1), Hu = make_halo(l2);
    double *  S = to_device(hS), *U = to_device(hU), *A = to_device(hA), *dU = to_device(hdU);  // T-vectors = prefixes
    double *  SL, *UL, *AL, *dUL;
    CUDA(cudaMalloc(&SL, sizeof(double) * size_t(l1.lsize)));
    CUDA(cudaMalloc(&UL, sizeof(double) * L2));
    CUDA(cudaMalloc(&AL, sizeof(double) * L2));
    CUDA(cudaMalloc(&dUL, sizeof(double) * L2));
    CHECK(sb_halo_P_nccl(Hs, comm, S, SL, nullptr));
    CHECK(sb_halo_P_nccl(Hu, comm, U, UL, nullptr));
    CHECK(sb_halo_P_nccl(Hu, comm, A, AL, nullptr));
    const double* in[3] = {SL, UL, AL};
    CHECK(sb_functional_evaluate(F, 0.0, 1, in, SB_DEVICE, r_L, SB_DEVICE, 0, nullptr));
    CHECK(sb_halo_Pt_nccl(Hu, comm, r_L, r_T, nullptr));
    CUDA(cudaDeviceSynchronize());
    write_bin(rd + "/out_nccl_r.bin", to_host(r_T, T2));
    CHECK(sb_halo_P_nccl(Hu, comm, dU, dUL, nullptr));
    CHECK(sb_functional_jvp(F, 1, dUL, SB_DEVICE, r_L, SB_DEVICE, nullptr));
    CHECK(sb_halo_Pt_nccl(Hu, comm, r_L, r_T, nullptr));
    CUDA(cudaDeviceSynchronize());
    write_bin(rd + "/out_nccl_dr.bin", to_host(r_T, T2));
    sb_halo_destroy(Hs);
    sb_halo_destroy(Hu);
  }
  // ===== transport 2: peer rings mapped with CUDA IPC, in-place exchanges =====
  {
    sb_halo_t Hs = make_halo(l1), Hu = make_halo(l2), Ha = make_halo(l2);  // one object per field in flight
    CHECK(sb_halo_enable_peer_ipc(Hs, comm));
    CHECK(sb_halo_enable_peer_ipc(Hu, comm));
    CHECK(sb_halo_enable_peer_ipc(Ha, comm));
    if (!sb_halo_owned_first(Hu)) {
      std::fprintf(stderr, "rank %d: partition is not numbered owned-first\n", rank);
      return 5;
    }
    double *S = to_device(hS), *U = to_device(hU), *A = to_device(hA), *dU = to_device(hdU);  // L-sized, T prefix
    for (int round = 0; round < 3; round++) {  // repeated rounds exercise the two-slot ring
      const sb_halo_t hs[3] = {Hs, Hu, Ha};
      double*         vs[3] = {S, U, A};
      CHECK(sb_halo_P_peer_inplace_multi(3, hs, vs, nullptr));
      const double* in[3] = {S, U, A};
      CHECK(sb_functional_evaluate(F, 0.0, 1, in, SB_DEVICE, r_L, SB_DEVICE, 0, nullptr));
      CHECK(sb_halo_Pt_peer_inplace(Hu, r_L, nullptr));
      CUDA(cudaDeviceSynchronize());
    }
    write_bin(rd + "/out_peer_r.bin", to_host(r_L, T2));
    const sb_halo_t hd[1] = {Hu};
    double*         vd[1] = {dU};
    CHECK(sb_halo_P_peer_inplace_multi(1, hd, vd, nullptr));
    CHECK(sb_functional_jvp(F, 1, dU, SB_DEVICE, r_L, SB_DEVICE, nullptr));
    CHECK(sb_halo_Pt_peer_inplace(Hu, r_L, nullptr));
    CUDA(cudaDeviceSynchronize());
    write_bin(rd + "/out_peer_dr.bin", to_host(r_L, T2));
    int st[3] = {0, 0, 0};
    CHECK(sb_halo_peer_status(Hs, &st[0]));
    CHECK(sb_halo_peer_status(Hu, &st[1]));
    CHECK(sb_halo_peer_status(Ha, &st[2]));
    if (st[0] || st[1] || st[2]) {
      std::fprintf(stderr, "rank %d: a wait for a neighbour timed out\n", rank);
      return 6;
    }
    CHECK(sb_comm_barrier(comm, nullptr));  // nobody unmaps a ring a neighbour may still write into
    sb_halo_destroy(Hs);
    sb_halo_destroy(Hu);
    sb_halo_destroy(Ha);
  }
  // ===== the Newton step's linear solve across the ranks: K = P^T A_L P in sub-assembled form, Jacobi, CG =====
  if (access((rd + "/ess_local.bin").c_str(), R_OK) == 0) {
    sb_halo_t Hs = make_halo(l1), Hu = make_halo(l2);
    double *  S = to_device(hS), *U = to_device(hU), *A = to_device(hA);
    double *  SL, *UL, *AL;
    CUDA(cudaMalloc(&SL, sizeof(double) * size_t(l1.lsize)));
    CUDA(cudaMalloc(&UL, sizeof(double) * L2));
    CUDA(cudaMalloc(&AL, sizeof(double) * L2));
    CHECK(sb_halo_P_nccl(Hs, comm, S, SL, nullptr));
    CHECK(sb_halo_P_nccl(Hu, comm, U, UL, nullptr));
    CHECK(sb_halo_P_nccl(Hu, comm, A, AL, nullptr));
    const double* in[3] = {SL, UL, AL};
    CHECK(sb_functional_evaluate(F, 0.0, 1, in, SB_DEVICE, r_L, SB_DEVICE, 0, nullptr));
    sb_csr_t pat = nullptr;
    CHECK(sb_csr_create(G, G, &pat));
    double* values = nullptr;
    CUDA(cudaMalloc(&values, sizeof(double) * size_t(sb_csr_nnz(pat))));
    CHECK(sb_functional_assemble(F, 1, &pat, values, SB_DEVICE, nullptr));  // rank-local: no communication
    auto         ess = read_bin<int>(rd + "/ess_local.bin");
    sb_par_csr_t K   = nullptr;
    CHECK(sb_par_csr_create(comm, Hu, pat, values, ess.data(), int64_t(ess.size()), &K));
    double *diag = nullptr, *dinv = nullptr, *x = nullptr;
    CUDA(cudaMalloc(&diag, sizeof(double) * T2));
    CUDA(cudaMalloc(&dinv, sizeof(double) * T2));
    CUDA(cudaMalloc(&x, sizeof(double) * T2));
    CHECK(sb_par_csr_diagonal(K, diag, nullptr));
    CHECK(sb_jacobi_setup(diag, dinv, int64_t(T2), nullptr));
    double* b = to_device(read_bin<double>(rd + "/b_T.bin"));
    int     converged = 0, iters = 0;
    double  final_norm = 0.0;
    CHECK(sb_par_cg_solve(K, dinv, b, x, 1e-12, 0.0, 5000, &converged, &iters, &final_norm, nullptr));
    CUDA(cudaDeviceSynchronize());
    write_bin(rd + "/out_cg_x.bin", to_host(x, T2));
    write_bin(rd + "/out_cg_info.bin", std::vector<double>{double(converged), double(iters), final_norm});
    // the operator itself on the right-hand side (compared with the global matrix)
    CHECK(sb_par_csr_mult(K, b, x, nullptr));
    CUDA(cudaDeviceSynchronize());
    write_bin(rd + "/out_Kb.bin", to_host(x, T2));
    sb_par_csr_destroy(K);
    sb_csr_destroy(pat);
    sb_halo_destroy(Hs);
    sb_halo_destroy(Hu);
  }
  // ===== a reduction over the ranks: the true dofs of the partitions add up to the global space =====
  {
    double* d = to_device(std::vector<double>{double(T2)});
    CHECK(sb_comm_allreduce_sum(comm, d, 1, nullptr));
    CUDA(cudaDeviceSynchronize());
    write_bin(rd + "/out_sum_tsize.bin", to_host(d, 1));
  }
  CHECK(sb_comm_barrier(comm, nullptr));
  sb_functional_destroy(F);
  sb_integral_destroy(I);
  sb_restriction_destroy(G);
  sb_restriction_destroy(G1);
  sb_comm_destroy(comm);
  std::printf("rank %d of %d ok\n", rank, world);
  return 0;
}

int main(int argc, char** argv)
{
  if (argc < 2) {
    std::fprintf(stderr, "usage: par_functional_test <directory>\n");
    return 2;
  }
  g_dir           = argv[1];
  const int world = int(read_meta(g_dir + "/meta.txt")["world"]);
  std::remove((g_dir + "/nccl_id.bin").c_str());
  std::vector<pid_t> kids;
  for (int r = 0; r < world; r++) {  // fork before any CUDA call: every rank gets its own context
    pid_t pid = fork();
    if (pid == 0) std::_Exit(run_rank(r, world));
    kids.push_back(pid);
  }
  int failures = 0;
  for (pid_t pid : kids) {
    int status = 0;
    waitpid(pid, &status, 0);
    if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) failures++;
  }
  return failures == 0 ? 0 : 1;
}
