"""Run under torchrun (one rank per GPU): the partitioned Functional (P -> kernels -> P^T with the shared-dof exchange)
against the same Functional evaluated on the whole mesh by one GPU.  Prints one 'par_check ok' line from rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/par_check.py [nx ny nz p]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

from serac_b200 import capi
from serac_b200.parallel import HaloExchange, ParFunctional
from serac_b200.structured import Partition, StructuredHexMesh, factor_ranks

PARAMS = [1.0, 100.0, 50.0, 1.0]


def build(dev, mesh, part, p, part1=None, order=None):
    """Functional for SolidMechanics NeoHookean on `part` (+ its p = 1 partition part1), or the whole mesh when None"""
    if part is not None:
        tab, tab1 = part.element_table(p, element_order=order), part1.element_table(1, element_order=order)
        nn, nn1, xyz = part.num_local_nodes, part1.num_local_nodes, part1.vertex_coordinates()
    else:
        tab, tab1 = mesh.element_table(p), mesh.element_table(1)
        nn, nn1, xyz = mesh.num_nodes(p), mesh.num_nodes(1), mesh.vertex_coordinates()
    G = capi.ElementRestriction(tab, 3, nn)
    G1 = capi.ElementRestriction(tab1, 3, nn1)
    nodes = torch.as_tensor(xyz, dtype=torch.float64, device=dev)
    X_q, J_q = capi.geometric_factors(capi.CUBE, 1, p + 1, G1, nodes)
    I = capi.Integral(f"solid_neohookean_hex_p{p}", PARAMS, X_q, J_q, tab.shape[0])
    F = capi.Functional(G, [G1, G, G])
    F.AddDomainIntegral(I, [0, 1, 2])
    return dict(F=F, G=G, G1=G1, I=I, keep=(nodes, X_q, J_q))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    nx, ny, nz, p = (int(a) for a in (sys.argv[1:5] if len(sys.argv) >= 5 else (8, 6, 4, 2)))
    mesh = StructuredHexMesh(nx, ny, nz, 2.0, 1.5, 1.0)
    ranks = factor_ranks(world)
    inplace = (os.environ.get("SB_HALO_TRANSPORT", "peer") == "peer" and os.environ.get("SB_HALO_OVERLAP", "1") != "0"
               and os.environ.get("SB_HALO_INPLACE", "1") != "0")
    part = Partition(mesh, ranks, rank, p, 3, owned_first=inplace)
    part1 = Partition(mesh, ranks, rank, 1, 3, owned_first=inplace)
    # measured: no gain at 2 GPUs, -2.5% at 8; its half-exchanges need the flag protocol (SB_HALO_LL=0)
    split = inplace and os.environ.get("SB_HALO_SPLIT", "0") != "0" and os.environ.get("SB_HALO_LL", "1") == "0"
    order, n_independent = part.halo_independent_first() if split else (None, 0)
    loc = build(dev, mesh, part, p, part1, order)
    halo1 = HaloExchange(part1.halo_lists(), dev)
    halo, halo_a, halo_du, halo_r = (HaloExchange(part.halo_lists(), dev) for _ in range(4))
    halos = [halo1, halo, halo_a, halo_du, halo_r]
    transport = "nccl"
    if os.environ.get("SB_HALO_TRANSPORT", "peer") == "peer":
        ok = [h.enable_peer() for h in halos]
        transport = "peer" if all(ok) else "nccl"
    overlap = transport == "peer" and os.environ.get("SB_HALO_OVERLAP", "1") != "0"
    if overlap:  # puts before waits, P(dU) prefetched and P^T(r) on a side stream (one halo object per field in flight)
        PF = ParFunctional(loc["F"], [halo1, halo, halo_a], halo, du_halos={1: halo_du}, out_halo=halo_r,
                           inplace=inplace, independent_elements=n_independent)
    else:
        PF = ParFunctional(loc["F"], [halo1, halo, halo], halo)
    # globally consistent inputs
    rng = np.random.default_rng(7)
    n, n1 = mesh.num_nodes(p) * 3, mesh.num_nodes(1) * 3
    U, A, S, dU = 0.01 * rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), 0.01 * rng.uniform(-1, 1, n1), rng.uniform(0, 1, n)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    gt, gt1 = part.global_true_vdofs(), part1.global_true_vdofs()
    def tvec(values, lsize):
        """T-vector; in-place mode: an L-sized buffer with the T-vector in its prefix (the rest is filled by P)"""
        v = t(values)
        if not inplace:
            return v
        buf = torch.full((lsize,), float("nan"), dtype=torch.float64, device=dev)
        buf[:len(values)] = v
        return buf

    Ts, Tu, Ta, Tdu = (tvec(S[gt1], loc["G1"].LSize()), tvec(U[gt], loc["G"].LSize()), tvec(A[gt], loc["G"].LSize()),
                       tvec(dU[gt], loc["G"].LSize()))
    for _ in range(6):  # repeated rounds exercise the two-slot ring
        if overlap:
            PF.prefetch_direction(1, Tdu)
        r_T = PF(0.0, [Ts, Tu, Ta], wrt=1)
        dr_T = PF.GradientMult(1, Tdu)
        PF.join()
    torch.cuda.synchronize()
    if transport == "peer":
        assert all(h.peer_status() == 0 for h in halos), "a peer flag wait timed out"
        transport += (("+overlap" if overlap else "") + ("+inplace" if inplace else "") +
                      ("+LL" if inplace and os.environ.get("SB_HALO_LL", "1") != "0" else "") +
                      (f"+split({n_independent}/{part.num_elements})" if split else ""))
    # the same on the whole mesh, one GPU
    glob = build(dev, mesh, None, p)
    r = glob["F"](0.0, [t(S), t(U), t(A)], wrt=1).cpu().numpy()
    dr = glob["F"].GradientMult(1, t(dU)).cpu().numpy()
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
    e1, e2 = rel(r_T.cpu().numpy(), r[gt]), rel(dr_T.cpu().numpy(), dr[gt])
    errs = torch.tensor([e1, e2], dtype=torch.float64, device=dev)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    ok = bool((errs < 1e-13).all())
    if rank == 0:
        print(f"par_check {'ok' if ok else 'FAILED'}: world={world} ranks={ranks} mesh={nx}x{ny}x{nz} p={p} transport={transport} "
              f"residual {errs[0].item():.2e} jvp {errs[1].item():.2e}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
