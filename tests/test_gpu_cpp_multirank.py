"""GPU (two devices): the multi-rank C ABI driven from C++ with no Python in the ranks -- NCCL bootstrap through
sb_comm_unique_id / sb_comm_create, P / P^T over ncclSend / ncclRecv (sb_halo_P_nccl / sb_halo_Pt_nccl) and over CUDA-IPC
peer rings (sb_halo_enable_peer_ipc + the in-place kernels) around sb_functional_evaluate / sb_functional_jvp
(tests/cpp/par_functional_test.cu).  pytest only writes the partition (serac_b200/structured.py, covered on the CPU by
tests/test_partition.py) and compares every rank's T-vectors with the same Functional evaluated on the whole mesh."""
import os
import subprocess

import numpy as np
import pytest

from harness import ROOT, rel_l2

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "tests", "cpp", "par_functional_test")
PARAMS = [1.0, 100.0, 50.0, 1.0]


def _export_halo(d, tag, lists):
    for k, dt in (("ltdof", np.int32), ("neighbors", np.int32), ("send_idx", np.int32), ("recv_idx", np.int32),
                  ("send_off", np.int64), ("recv_off", np.int64)):
        np.ascontiguousarray(lists[k], dtype=dt).tofile(os.path.join(d, f"{tag}_{k}.bin"))


def test_two_ranks_from_cpp_match_the_whole_mesh(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    if not os.path.exists(BIN):
        pytest.skip("tests/cpp/par_functional_test not built (make -C tests/cpp)")
    from serac_b200 import capi
    from serac_b200.structured import Partition, StructuredHexMesh
    world, ranks, p = 2, (2, 1, 1), 2
    mesh = StructuredHexMesh(6, 4, 3, 2.0, 1.5, 1.0)
    n, n1 = mesh.num_nodes(p) * 3, mesh.num_nodes(1) * 3
    rng = np.random.default_rng(7)
    U, A, S, dU = 0.01 * rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), 0.01 * rng.uniform(-1, 1, n1), rng.uniform(0, 1, n)
    # the whole mesh on one GPU
    dev = torch.device("cuda:0")
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    G = capi.ElementRestriction(mesh.element_table(p), 3, mesh.num_nodes(p))
    G1 = capi.ElementRestriction(mesh.element_table(1), 3, mesh.num_nodes(1))
    X_q, J_q = capi.geometric_factors(capi.CUBE, 1, p + 1, G1, t(mesh.vertex_coordinates()))
    I = capi.Integral("solid_neohookean_hex_p2", PARAMS, X_q, J_q, mesh.ne)
    F = capi.Functional(G, [G1, G, G])
    F.AddDomainIntegral(I, [0, 1, 2])
    r = F(0.0, [t(S), t(U), t(A)], wrt=1).cpu().numpy()
    dr = F.GradientMult(1, t(dU)).cpu().numpy()
    # the Newton step's linear solve on the whole mesh: Jacobian eliminated at the face x = 0, Jacobi, CG
    gx = mesh.node_grid(p)[0]
    fixed = np.nonzero((np.arange(n) // 3) % gx == 0)[0].astype(np.int32)
    b = rng.uniform(-1, 1, n)
    b[fixed] = 0.0
    pat, vals = F.assemble(1)
    import scipy.sparse as sp
    capi.eliminate_rows_cols(pat, vals, fixed, want_Ae=False)
    rp, ci = pat.to_host()
    K_global = sp.csr_matrix((vals.cpu().numpy(), ci, rp), shape=(n, n))
    dinv = capi.jacobi_setup(capi.csr_diagonal(pat, vals))
    cg = capi.CGSolver(n, rel_tol=1e-12, max_iter=5000)
    x_global, info_global = cg.Mult(pat, t(b), values=vals, dinv=dinv)
    assert info_global["converged"]
    x_global = x_global.cpu().numpy()
    # the partition, one directory per rank
    d = str(tmp_path)
    with open(os.path.join(d, "meta.txt"), "w") as f:
        f.write(f"world {world}\n")
    parts = []
    for rk in range(world):
        part = Partition(mesh, ranks, rk, p, 3, owned_first=True)
        part1 = Partition(mesh, ranks, rk, 1, 3, owned_first=True)
        parts.append(part)
        rd = os.path.join(d, f"rank{rk}")
        os.makedirs(rd)
        tab, tab1 = part.element_table(p), part1.element_table(1)
        with open(os.path.join(rd, "meta.txt"), "w") as f:
            f.write(f"ne {tab.shape[0]}\nnum_nodes {part.num_local_nodes}\nnum_nodes1 {part1.num_local_nodes}\n")
        np.ascontiguousarray(tab, dtype=np.uint64).tofile(os.path.join(rd, "tab.bin"))
        np.ascontiguousarray(tab1, dtype=np.uint64).tofile(os.path.join(rd, "tab1.bin"))
        np.ascontiguousarray(part1.vertex_coordinates(), dtype=np.float64).tofile(os.path.join(rd, "coords.bin"))
        l2, l1 = part.halo_lists(), part1.halo_lists()
        _export_halo(rd, "p2", l2)
        _export_halo(rd, "p1", l1)
        gt, gt1 = part.global_true_vdofs(), part1.global_true_vdofs()
        np.nonzero(np.isin(part.global_local_vdofs(), fixed))[0].astype(np.int32).tofile(os.path.join(rd, "ess_local.bin"))
        b[gt].tofile(os.path.join(rd, "b_T.bin"))
        for name, vals, lsize in (("U", U[gt], l2["lsize"]), ("A", A[gt], l2["lsize"]), ("dU", dU[gt], l2["lsize"]),
                                  ("S", S[gt1], l1["lsize"])):
            buf = np.full(lsize, np.nan)  # L-sized; the prefix is the T-vector, the rest is filled by P
            buf[:len(vals)] = vals
            buf.tofile(os.path.join(rd, name + ".bin"))
    env = dict(os.environ)
    out = subprocess.run([BIN, d], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    total = 0.0
    for rk, part in enumerate(parts):
        rd = os.path.join(d, f"rank{rk}")
        gt = part.global_true_vdofs()
        for tag in ("nccl", "peer"):
            r_T = np.fromfile(os.path.join(rd, f"out_{tag}_r.bin"))
            dr_T = np.fromfile(os.path.join(rd, f"out_{tag}_dr.bin"))
            assert rel_l2(r_T, r[gt]) < 1e-13 and rel_l2(dr_T, dr[gt]) < 1e-13, (rk, tag)
        # K = P^T A_L P (sub-assembled, include/serac_b200_comm.h) acts and solves like the global eliminated matrix
        assert rel_l2(np.fromfile(os.path.join(rd, "out_Kb.bin")), (K_global @ b)[gt]) < 1e-13, rk
        conv, iters, _ = np.fromfile(os.path.join(rd, "out_cg_info.bin"))
        assert conv == 1.0 and abs(int(iters) - info_global["final_iter"]) <= 2, (iters, info_global)
        assert rel_l2(np.fromfile(os.path.join(rd, "out_cg_x.bin")), x_global[gt]) < 1e-9, rk
        total = float(np.fromfile(os.path.join(rd, "out_sum_tsize.bin"))[0])
        assert total == n  # ncclAllReduce over the ranks: the true dofs add up to the global space
