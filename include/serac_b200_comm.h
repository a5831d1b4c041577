/* serac_b200 C ABI, multi-rank part: one process per GPU, the ranks of one mesh partition.
 *
 * Replaces, for the hot path, what the reference gets from MPI through mfem: the shared-dof exchange of
 * mfem::ConformingProlongationOperator (mfem/fem/pfespace.cpp:3599-3663; its device variant :3740-3933 posts
 * MPI_Isend / MPI_Irecv per neighbour) applied around every Functional evaluation (F/functional.hpp:452-454, :479), and
 * the MPI_Allreduce of quantities of interest (F/functional_qoi.inl).  Transports:
 *   sb_halo_P_nccl / sb_halo_Pt_nccl      pack kernel -> ncclGroupStart, ncclSend / ncclRecv per neighbour, ncclGroupEnd ->
 *                                         ordered unpack kernel, all on the caller's stream (NCCL over NVLink / NVSwitch)
 *   sb_halo_enable_peer_ipc + sb_halo_*_peer*   the neighbours' receive rings mapped with CUDA IPC; kernels put straight
 *                                         into them (include/serac_b200.h, "Peer-memory transport")
 * NCCL is loaded at run time (dlopen of libnccl.so.2, or the path in SB_NCCL_LIB): the library has no link-time dependency
 * on it and single-rank users never touch it.  Bootstrap follows NCCL's own model: rank 0 calls sb_comm_unique_id and
 * distributes the 128 bytes out of band (a file, an environment variable, MPI_Bcast in an MPI program), then every rank
 * calls sb_comm_create on the device it has made current.
 */
#ifndef SERAC_B200_COMM_H
#define SERAC_B200_COMM_H

#include "serac_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sb_comm_s* sb_comm_t;
#define SB_COMM_ID_BYTES 128

int sb_comm_unique_id(void* id_bytes);                                           /* ncclGetUniqueId */
int sb_comm_create(const void* id_bytes, int rank, int world, sb_comm_t* out);   /* ncclCommInitRank, current device */
int sb_comm_destroy(sb_comm_t c);
int sb_comm_rank(sb_comm_t c);
int sb_comm_size(sb_comm_t c);
/* sum over the ranks in place (quantities of interest, dot products of a distributed Krylov solver) */
int sb_comm_allreduce_sum(sb_comm_t c, double* values_dev, int64_t n, void* stream);
/* device-side barrier of the ranks followed by a host synchronisation of `stream` */
int sb_comm_barrier(sb_comm_t c, void* stream);

/* ranks of the halo object's neighbours, in the order of its send / receive lists (sb_halo_create) */
int sb_halo_set_neighbor_ranks(sb_halo_t H, const int* neighbor_ranks);
/* P: T -> L and P^T: L -> T over NCCL point-to-point (fixed neighbour order in the unpack: deterministic sums) */
int sb_halo_P_nccl(sb_halo_t H, sb_comm_t c, const double* T_dev, double* L_dev, void* stream);
int sb_halo_Pt_nccl(sb_halo_t H, sb_comm_t c, const double* L_dev, double* T_dev, void* stream);
/* Collective over the communicator: allocates this rank's receive ring, exchanges CUDA IPC handles and the neighbour
 * tables (ncclAllGather), maps the neighbours' rings and enables the peer-memory transport of H (sb_halo_*_peer*).
 * Every rank calls it the same number of times, for halo objects built from the same partition. */
int sb_halo_enable_peer_ipc(sb_halo_t H, sb_comm_t c);

/* ---------------------------------------------------------------------------------------------------------------
 * The assembled Jacobian of a partitioned Functional as an operator on T-vectors, and mfem::CGSolver on it.
 *
 * The reference turns the rank-local matrix A_L into K = P^T A_L P with hypre's RAP (F/functional.hpp:654-662) and hands K
 * to hypre's solvers.  Without hypre the product stays in its factors (sub-assembled form): every rank keeps the CSR
 * matrix of its own elements over its L-dofs (what sb_functional_assemble writes, no communication) and
 * y_T = P^T (A_L (P x_T)), with the exchanges that surround every evaluation.  diag K = P^T diag A_L (Jacobi).
 * sb_par_csr_create eliminates the essential dofs (local L ids, every local copy, owned or not) from the local values IN
 * PLACE -- rows and columns zeroed, unit diagonal on the owner's copy only, so that K = [K_ii 0; 0 I] as
 * eliminateAllEssentialDofsFromMatrix leaves it (boundary_condition_manager.hpp:91-94).  H needs its neighbour ranks
 * (sb_halo_set_neighbor_ranks); the exchanges run over NCCL point-to-point.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct sb_par_csr_s* sb_par_csr_t;
int sb_par_csr_create(sb_comm_t c, sb_halo_t H, sb_csr_t A_L, double* values_dev, const int* ess_local_dofs_host,
                      int64_t num_ess, sb_par_csr_t* out);
int sb_par_csr_destroy(sb_par_csr_t K);
int sb_par_csr_mult(sb_par_csr_t K, const double* x_T_dev, double* y_T_dev, void* stream);
int sb_par_csr_diagonal(sb_par_csr_t K, double* diag_T_dev, void* stream);
/* (x, y) over the true dofs of all ranks (every true dof lives on exactly one rank); synchronises the stream */
int sb_par_dot(sb_comm_t c, const double* x_T_dev, const double* y_T_dev, int64_t tsize, double* result_host, void* stream);
/* mfem::CGSolver (mfem/linalg/solvers.cpp:719-897) on T-vectors across the ranks: the recurrences and the stopping test
 * of sb_cg_solve_csr with every inner product reduced over the ranks (two reductions per iteration, as hypre's PCG).
 * x_T_dev is the result (initial guess 0); diag_inv_T_dev = Jacobi scaling or NULL. */
int sb_par_cg_solve(sb_par_csr_t K, const double* diag_inv_T_dev, const double* b_T_dev, double* x_T_dev, double rel_tol,
                    double abs_tol, int max_iter, int* converged, int* final_iter, double* final_norm, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SERAC_B200_COMM_H */
