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

#ifdef __cplusplus
}
#endif
#endif /* SERAC_B200_COMM_H */
