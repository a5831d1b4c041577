"""Run under torchrun (one rank per GPU): a quasi-static Newton solve on a PARTITIONED mesh -- residual through
ParFunctional, Jacobian assembled rank-locally and applied in sub-assembled form y = P^T A_L P x (ParCsrOperator), Jacobi
from P^T diag(A_L), CG with inner products reduced over the ranks (par_cg) -- against the same solve on the whole mesh
on one GPU with the rank-local solver layer (sb_cg_*).  The reference does this step with hypre: RAP of the local matrix
(functional.hpp:654-662), eliminateAllEssentialDofsFromMatrix, HyprePCG / HypreGMRES.  Prints one line from rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
        tools/par_newton_check.py [nx ny nz]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

from serac_b200 import capi
from serac_b200.parallel import HaloExchange, ParCsrOperator, ParFunctional, par_cg, par_dot
from serac_b200.structured import Partition, StructuredHexMesh, factor_ranks

PARAMS = [1.0, 100.0, 50.0, 1.0]
P_ORDER = 2
REL_TOL, NEWTON_REL, NEWTON_MAX = 1e-11, 1e-10, 12


def functional(dev, tab, tab1, nn, nn1, xyz):
    G = capi.ElementRestriction(tab, 3, nn)
    G1 = capi.ElementRestriction(tab1, 3, nn1)
    nodes = torch.as_tensor(xyz, dtype=torch.float64, device=dev)
    X_q, J_q = capi.geometric_factors(capi.CUBE, 1, P_ORDER + 1, G1, nodes)
    I = capi.Integral(f"solid_neohookean_hex_p{P_ORDER}", PARAMS, X_q, J_q, tab.shape[0])
    F = capi.Functional(G, [G1, G, G])
    F.AddDomainIntegral(I, [0, 1, 2])
    return dict(F=F, G=G, G1=G1, keep=(I, nodes, X_q, J_q))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    nx, ny, nz = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (8, 4, 4)))
    mesh = StructuredHexMesh(nx, ny, nz, 2.0, 1.0, 1.0)
    ranks = factor_ranks(world)
    p = P_ORDER
    t64 = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    gx = mesh.node_grid(p)[0]
    n, n1 = mesh.num_nodes(p) * 3, mesh.num_nodes(1) * 3
    gid = np.arange(n)
    body = 2.0 * np.sin(0.7 * gid + 0.3)  # "acceleration" field: a body force through the mass term
    fixed = np.nonzero((gid // 3) % gx == 0)[0]  # all components of the nodes on the face x = 0

    # ---- the partitioned solve ----
    part, part1 = Partition(mesh, ranks, rank, p, 3), Partition(mesh, ranks, rank, 1, 3)
    loc = functional(dev, part.element_table(p), part1.element_table(1), part.num_local_nodes, part1.num_local_nodes,
                     part1.vertex_coordinates())
    halo, halo1 = HaloExchange(part.halo_lists(), dev), HaloExchange(part1.halo_lists(), dev)
    PF = ParFunctional(loc["F"], [halo1, halo, halo], halo)
    gt, gl = part.global_true_vdofs(), part.global_local_vdofs()
    ess_L = np.nonzero(np.isin(gl, fixed))[0].astype(np.int32)
    ess_T = torch.as_tensor(np.nonzero(np.isin(gt, fixed))[0].astype(np.int64), device=dev)
    S_T = torch.zeros(halo1.tsize, dtype=torch.float64, device=dev)
    a_T, u_T = t64(body[gt]), torch.zeros(halo.tsize, dtype=torch.float64, device=dev)
    norms, its = [], []
    for _ in range(NEWTON_MAX):
        r_T = PF(0.0, [S_T, u_T, a_T], wrt=1)
        r_T[ess_T] = 0.0
        norms.append(np.sqrt(par_dot(r_T, r_T)))
        if norms[-1] < NEWTON_REL * norms[0]:
            break
        pat, vals = loc["F"].assemble(1)  # rank-local, no communication
        A = ParCsrOperator(pat, vals, halo, ess_L)
        du, info = par_cg(A, r_T, 1.0 / A.diagonal(), rel_tol=REL_TOL, max_iter=5000)
        assert info["converged"], info
        its.append(info["final_iter"])
        u_T = u_T - du

    # ---- the same on the whole mesh, one GPU, rank-local solver layer ----
    u_g = torch.zeros(n, dtype=torch.float64, device=dev)
    ref_norms, ref_its = [], []
    if rank == 0:
        glob = functional(dev, mesh.element_table(p), mesh.element_table(1), mesh.num_nodes(p), mesh.num_nodes(1),
                          mesh.vertex_coordinates())
        F = glob["F"]
        S, a = torch.zeros(n1, dtype=torch.float64, device=dev), t64(body)
        ess = fixed.astype(np.int32)
        cg = capi.CGSolver(n, rel_tol=REL_TOL, max_iter=5000)
        for _ in range(NEWTON_MAX):
            r = F(0.0, [S, u_g, a], wrt=1)
            capi.set_subvector(r, ess, 0.0)
            ref_norms.append(float(torch.linalg.norm(r)))
            if ref_norms[-1] < NEWTON_REL * ref_norms[0]:
                break
            pat, vals = F.assemble(1)
            capi.eliminate_rows_cols(pat, vals, ess, want_Ae=False)
            dinv = capi.jacobi_setup(capi.csr_diagonal(pat, vals))
            du, info = cg.Mult(pat, r, values=vals, dinv=dinv)
            assert info["converged"], info
            ref_its.append(info["final_iter"])
            u_g -= du
    dist.broadcast(u_g, 0)
    gt_t = torch.as_tensor(gt, dtype=torch.int64, device=dev)
    acc = torch.stack([((u_T - u_g[gt_t])**2).sum(), (u_g[gt_t]**2).sum()])
    dist.all_reduce(acc)
    err = float((acc[0] / acc[1]).sqrt().item())
    meta = [None]
    if rank == 0:
        meta = [dict(norms=ref_norms, its=ref_its)]
    dist.broadcast_object_list(meta, 0)
    ok = (err < 1e-8 and len(norms) == len(meta[0]["norms"]) and
          all(abs(i - j) <= 2 for i, j in zip(its, meta[0]["its"])) and norms[-1] < NEWTON_REL * norms[0])
    if rank == 0:
        print(f"par_newton_check {'ok' if ok else 'FAILED'}: world={world} ranks={ranks} mesh={nx}x{ny}x{nz} p={p} "
              f"|u - u_global| / |u_global| = {err:.2e}; Newton residuals {['%.2e' % v for v in norms]} vs "
              f"{['%.2e' % v for v in meta[0]['norms']]}; CG iterations {its} vs {meta[0]['its']}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
