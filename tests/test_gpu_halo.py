"""GPU: the shared-dof exchange kernels (sb_halo_*, capi.cu "Shared-dof exchange") on ONE device: the ranks of a box
partition live side by side in this process, each with its own halo object, staging buffers, receive ring and stream --
rank r's neighbour pointers address rank s's ring directly, which is exactly what the IPC / symmetric-memory mapping
gives on several GPUs.  Every transport is exercised: pack / finish around a copied transport (the NCCL send/recv path),
the put + flag protocol, and the LL protocol whose kernels put, wait and unpack in one launch (ranks on concurrent
streams).  Expected values come from the global vector directly (P copies, P^T sums: mfem/fem/pfespace.cpp:3599-3663).
A second test runs tools/par_check.py on two real ranks when two GPUs are visible."""
import os
import subprocess
import sys

import numpy as np
import pytest

from harness import ROOT

pytestmark = pytest.mark.gpu


def _setup(ranks, p=2, mesh_n=(4, 3, 2)):
    import torch
    from serac_b200.parallel import HaloExchange
    from serac_b200.structured import Partition, StructuredHexMesh
    dev = torch.device("cuda:0")
    mesh = StructuredHexMesh(*mesh_n)
    world = int(np.prod(ranks))
    parts = [Partition(mesh, ranks, r, p, 3, owned_first=True) for r in range(world)]
    lists = [q.halo_lists() for q in parts]
    nglob = mesh.num_nodes(p) * 3
    return torch, HaloExchange, dev, parts, lists, nglob, world


def _expected_Pt(parts, Ls, nglob):
    tot = np.zeros(nglob)
    for q, L in zip(parts, Ls):
        np.add.at(tot, q.global_local_vdofs(), L)
    return [tot[q.global_true_vdofs()] for q in parts]


def _attach(torch, HaloExchange, halos, lists, dev):
    slot = max(max(int(h.send_off[-1]), int(h.recv_off[-1]), 1) for h in halos)
    rings = [torch.zeros(HaloExchange.ring_doubles(slot), dtype=torch.float64, device=dev) for _ in halos]
    meta = [dict(neighbors=l["neighbors"], send_off=np.asarray(l["send_off"]).tolist(),
                 recv_off=np.asarray(l["recv_off"]).tolist()) for l in lists]
    for r, h in enumerate(halos):
        h.attach_rings(rings[r], slot, r, [x.data_ptr() for x in rings], meta)
    return rings


@pytest.mark.parametrize("ranks", [(2, 1, 1), (2, 2, 1)])
def test_pack_finish_with_copied_transport(ranks):
    """sb_halo_P_pack / P_finish / Pt_pack / Pt_finish: what runs around ncclSend / ncclRecv"""
    torch, HaloExchange, dev, parts, lists, nglob, world = _setup(ranks)
    from serac_b200 import capi
    lib = capi.lib()
    halos = [HaloExchange(l, dev) for l in lists]
    rng = np.random.default_rng(3)
    g = rng.uniform(-1, 1, nglob)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    s = capi._stream_ptr()

    def transport(out_key, in_key):  # neighbour i of rank r receives segment i of r's send buffer
        for r, h in enumerate(halos):
            for i, n in enumerate(h.neighbors):
                j = halos[n].neighbors.index(r)
                src = h.send_buf[getattr(h, out_key)[i]:getattr(h, out_key)[i + 1]]
                dst = halos[n].recv_buf[getattr(halos[n], in_key)[j]:getattr(halos[n], in_key)[j + 1]]
                dst.copy_(src)

    Ts = [t(g[q.global_true_vdofs()]) for q in parts]
    for h, T in zip(halos, Ts):
        capi.check(lib.sb_halo_P_pack(h.h, capi._dev_ptr(T), s))
    transport("send_off", "recv_off")
    for q, h, T in zip(parts, halos, Ts):
        L = torch.full((h.lsize,), float("nan"), dtype=torch.float64, device=dev)
        capi.check(lib.sb_halo_P_finish(h.h, capi._dev_ptr(T), capi._dev_ptr(L), s))
        assert np.array_equal(L.cpu().numpy(), g[q.global_local_vdofs()])
    Ls = [rng.uniform(-1, 1, h.lsize) for h in halos]
    Ld = [t(L) for L in Ls]
    for h, L in zip(halos, Ld):
        capi.check(lib.sb_halo_Pt_pack(h.h, capi._dev_ptr(L), s))
    transport("recv_off", "send_off")
    for h, L, want in zip(halos, Ld, _expected_Pt(parts, Ls, nglob)):
        T = torch.full((h.tsize,), float("nan"), dtype=torch.float64, device=dev)
        capi.check(lib.sb_halo_Pt_finish(h.h, capi._dev_ptr(L), capi._dev_ptr(T), s))
        assert np.allclose(T.cpu().numpy(), want, rtol=0, atol=1e-14)


@pytest.mark.parametrize("ranks", [(2, 1, 1), (2, 2, 1)])
def test_flag_protocol_put_then_finish(ranks):
    """sb_halo_P_peer_put / _finish, sb_halo_Pt_peer_put / _finish over several rounds (the two-slot ring)"""
    torch, HaloExchange, dev, parts, lists, nglob, world = _setup(ranks)
    halos = [HaloExchange(l, dev) for l in lists]
    rings = _attach(torch, HaloExchange, halos, lists, dev)
    from serac_b200 import capi
    lib = capi.lib()
    rng = np.random.default_rng(5)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    s = capi._stream_ptr()
    for rnd in range(5):
        g = rng.uniform(-1, 1, nglob)
        Ts = [t(g[q.global_true_vdofs()]) for q in parts]
        for h, T in zip(halos, Ts):
            h.P_put(T)
        for q, h, T in zip(parts, halos, Ts):
            L = h.P_finish(T)
            assert np.array_equal(L.cpu().numpy(), g[q.global_local_vdofs()]), rnd
        Ls = [rng.uniform(-1, 1, h.lsize) for h in halos]
        Ld = [t(L) for L in Ls]
        for h, L in zip(halos, Ld):
            capi.check(lib.sb_halo_Pt_peer_put(h.h, capi._dev_ptr(L), s))
        for h, L, want in zip(halos, Ld, _expected_Pt(parts, Ls, nglob)):
            T = torch.full((h.tsize,), float("nan"), dtype=torch.float64, device=dev)
            capi.check(lib.sb_halo_Pt_peer_finish(h.h, capi._dev_ptr(L), capi._dev_ptr(T), s))
            assert np.allclose(T.cpu().numpy(), want, rtol=0, atol=1e-14), rnd
    torch.cuda.synchronize()
    assert all(h.peer_status() == 0 for h in halos)
    del rings


@pytest.mark.parametrize("ranks", [(2, 1, 1), (2, 2, 2)])
def test_inplace_exchanges_ranks_on_concurrent_streams(ranks):
    """sb_halo_P_peer_inplace_multi (two fields in one launch) and sb_halo_Pt_peer_inplace: the default (LL) protocol puts,
    waits and unpacks in ONE kernel per rank, so the ranks' kernels must run side by side -- one stream per rank"""
    torch, HaloExchange, dev, parts, lists, nglob, world = _setup(ranks, mesh_n=(4, 4, 2))
    import ctypes as C
    from serac_b200 import capi
    lib = capi.lib()
    # two fields in flight per rank -> two halo objects per rank, each set with its own rings
    halos_a = [HaloExchange(l, dev) for l in lists]
    halos_b = [HaloExchange(l, dev) for l in lists]
    rings = [_attach(torch, HaloExchange, hs, lists, dev) for hs in (halos_a, halos_b)]
    assert all(h.owned_first() for h in halos_a)
    streams = [torch.cuda.Stream() for _ in range(world)]
    rng = np.random.default_rng(11)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    for rnd in range(4):
        ga, gb = rng.uniform(-1, 1, nglob), rng.uniform(-1, 1, nglob)
        bufs = []
        for q, ha in zip(parts, halos_a):  # L-sized buffers whose prefix is the T-vector; the rest is filled by P
            ba = torch.full((ha.lsize,), float("nan"), dtype=torch.float64, device=dev)
            bb = torch.full((ha.lsize,), float("nan"), dtype=torch.float64, device=dev)
            ba[:ha.tsize] = t(ga[q.global_true_vdofs()])
            bb[:ha.tsize] = t(gb[q.global_true_vdofs()])
            bufs.append((ba, bb))
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                hs = (C.c_void_p * 2)(halos_a[r].h, halos_b[r].h)
                vs = (C.c_void_p * 2)(bufs[r][0].data_ptr(), bufs[r][1].data_ptr())
                capi.check(lib.sb_halo_P_peer_inplace_multi(2, hs, vs, C.c_void_p(streams[r].cuda_stream)))
        torch.cuda.synchronize()
        for q, (ba, bb) in zip(parts, bufs):
            assert np.array_equal(ba.cpu().numpy(), ga[q.global_local_vdofs()]), rnd
            assert np.array_equal(bb.cpu().numpy(), gb[q.global_local_vdofs()]), rnd
        Ls = [rng.uniform(-1, 1, h.lsize) for h in halos_a]
        Ld = [t(L) for L in Ls]
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                capi.check(lib.sb_halo_Pt_peer_inplace(halos_a[r].h, capi._dev_ptr(Ld[r]),
                                                       C.c_void_p(streams[r].cuda_stream)))
        torch.cuda.synchronize()
        for h, L, want in zip(halos_a, Ld, _expected_Pt(parts, Ls, nglob)):
            assert np.allclose(L[:h.tsize].cpu().numpy(), want, rtol=0, atol=1e-14), rnd
    assert all(h.peer_status() == 0 for h in halos_a + halos_b)
    del rings


def test_partitioned_functional_on_two_gpus_matches_global():
    """tools/par_check.py on two ranks (NCCL bootstrap, symmetric-memory rings, P -> kernels -> P^T) against the same
    Functional on the whole mesh; skipped on single-GPU machines"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "par_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "par_check ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_partitioned_newton_solve_on_two_gpus_matches_global():
    """tools/par_newton_check.py: the Newton step's linear algebra at N ranks (Jacobian in sub-assembled form
    y = P^T A_L P x, Jacobi from P^T diag(A_L), CG with rank-reduced inner products) against the single-GPU solve"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29534", os.path.join(ROOT, "tools", "par_newton_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "par_newton_check ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
