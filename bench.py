#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: GDOF/s of the Functional residual + Jacobian action, p=2 hex NeoHookean.

One *step* = what a Newton-Krylov iteration asks of serac::Functional on config A (BASELINE.json configs[1]:
SolidMechanics NeoHookean, H1 p=2 hex, beam-hex.mesh refined 4x = 32 768 hexes / 839 619 vdofs / 884 736 qpts):

    auto [r, drdu] = residual(t, shape, differentiate_wrt(u), a);   // residual + stored dqf   (functional.hpp:446-498)
    drdu.Mult(du, dr);                                              // Jacobian action          (functional.hpp:406-432)

i.e. gather -> evaluation kernel with derivative store -> deterministic scatter, then gather -> JVP kernel -> scatter.
`value`  : vdofs / step time with all inputs resident in HBM (CUDA-graph replay of the step, CUDA events).
`e2e`    : the same step through the host-buffer C ABI (sb_functional_evaluate / sb_functional_jvp with pinned host
           L-vectors in, host L-vectors out; the copies are inside the timed region).
`roofline`: the Jacobian-action kernel (the pure HBM stream over the stored derivatives) timed with CUDA events.
`--impl reference`: the REFERENCE's own CPU kernels (oracle/_ref) on the host cores, same step, one mesh slab per core.

Launch: python bench.py [--gpus N --steps K --warmup W]; for N > 1 under torchrun (one rank per GPU)."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GDOF/s Functional residual + Jacobian action (p=2 hex)"
CONFIG_A = os.path.join(ROOT, "tests", "golden", "config_a_beam_hex_r4_p2.npz")
NEOHOOKEAN = [1.0, 100.0, 50.0, 1.0]  # density, K, G, geometric nonlinearity on (BASELINE.md §2 probe values)
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this command
# (profiles/r4_summary.md, raw extract profiles/r4b_ncu_extract.txt; symmetric tangent storage, 45 doubles per point)
NCU_TRAFFIC_SOURCE = ("ncu --set full capture of this command, one launch, not measured in this run: "
                      "profiles/r4b_ncu_extract.txt (dram__bytes_read.sum + dram__bytes_write.sum)")
JVP_NCU_TRAFFIC_BYTES = 348.7e6   # 328.8 MB read + 19.9 MB written
EVAL_NCU_TRAFFIC_BYTES = 363.8e6  # 82.9 MB read + 280.9 MB written (rest of the 319 MB store still in L2)
# what the same capture says limits the two kernels now that they move 44% fewer bytes
JVP_LIMITER = ("shared-memory data pipe ~70% of cycles (516 wavefronts per element before the scratch re-layout), "
               "HBM 48%, FP64 16% (profiles/r4b_ncu_extract.txt)")
EVAL_LIMITER = ("no single pipe: FP64 40%, shared-memory + global wavefronts ~64% of cycles, issue slots 42%, 16 warps/SM "
                "at 128 registers (profiles/r4b_ncu_extract.txt)")


# ---------------------------------------------------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi samples of SM clock / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.samples.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
def build_problem(device):
    """config A on one GPU: product objects built from the host library's native dofs (fixture written by
    oracle/make_golden.py from mfem: the mesh / fespace layer is out of scope and stays on the host)."""
    import torch
    from serac_b200 import capi
    d = np.load(CONFIG_A)
    geom, p, dim = int(d["geom"]), int(d["p"]), int(d["dim"])
    info = capi.h1_dof_info(geom, p, d["native_dofs"])
    shape_info = capi.h1_dof_info(geom, 1, d["shape_native_dofs"])
    G = capi.ElementRestriction(info, dim, int(d["num_nodes"]))
    G_shape = capi.ElementRestriction(shape_info, dim, int(d["shape_num_nodes"]))
    G_mesh = capi.ElementRestriction(shape_info, dim, int(d["mesh_num_nodes"]), int(d["mesh_ordering"]))
    nodes = torch.as_tensor(d["mesh_nodes"], dtype=torch.float64, device=device)
    X_q, J_q = capi.geometric_factors(geom, int(d["mesh_order"]), p + 1, G_mesh, nodes)
    integral = capi.Integral("solid_neohookean_hex_p2", NEOHOOKEAN, X_q, J_q, int(d["ne"]))
    F = capi.Functional(G, [G_shape, G, G])
    F.AddDomainIntegral(integral, [0, 1, 2])
    rng = np.random.default_rng(42)
    vs, ss = G.LSize(), G_shape.LSize()
    host = dict(U=1e-3 * rng.uniform(-1, 1, vs), A=np.zeros(vs), S=np.zeros(ss), dU=rng.uniform(0, 1, vs))
    return dict(capi=capi, G=G, G_shape=G_shape, G_mesh=G_mesh, integral=integral, F=F, host=host, ne=int(d["ne"]),
                nq=int(d["ne"]) * (p + 1)**dim, vdofs=vs, tdofs=vs, tdofs_shape=ss, p=p, keep=(nodes, X_q, J_q),
                mesh="beam-hex.mesh refined 4x (mfem numbering; fixture tests/golden/config_a_beam_hex_r4_p2.npz)")


# Workload selection.  Default (what the driver runs, BASELINE.json's metric): config A = configs[1].
# SB_BENCH_CONFIG=b selects configs[4]: synthetic structured hex cube, H1 p=3, NeoHookean --
#   weak scaling (default):  one cube of SB_BENCH_B_N^3 hexes per GPU (default 107: ~100 M vdofs and 51 GB of stored
#                            derivatives per GPU; 135 = ~200 M vdofs per GPU), boxes laid out 2x1x1 / 2x2x1 / 2x2x2
#   SB_BENCH_SCALING=strong: the fixed 135^3 cube (200.8 M vdofs) split over the GPUs
BENCH_CONFIG = os.environ.get("SB_BENCH_CONFIG", "a").lower()
STRONG_SCALING = os.environ.get("SB_BENCH_SCALING", "weak") == "strong"
if BENCH_CONFIG == "b":
    ORDER = 3
    _bn = int(os.environ.get("SB_BENCH_B_N", "107"))
    PER_GPU_BOX, STRONG_BOX, ELEMENT_SIZE = (_bn, _bn, _bn), (135, 135, 135), 1.0 / 128.0
    WORKLOAD = (f"configs[4]: synthetic structured hex cube, H1 p=3, SolidMechanics NeoHookean, {_bn}^3 hexes "
                f"(~{3 * (3 * _bn + 1)**3 / 1e6:.0f} M vdofs) per GPU; step = residual with derivative store + one Jacobian action")
    STRONG_WORKLOAD = ("configs[4] strong scaling: the fixed 135^3-hex cube (200.8 M vdofs), H1 p=3 NeoHookean, split over "
                       "the GPUs; step = residual with derivative store + one Jacobian action")
else:
    ORDER = 2
    PER_GPU_BOX = (128, 16, 16)  # hexes per GPU = config A's 32 768 (the beam refined 4x), element size 1/16
    # optional strong-scaling series (N > 1 only; the default and the driver's series are weak): one fixed 128 x 32 x 32
    # mesh (131 072 hexes, 3.3 M vdofs) split over the GPUs
    STRONG_BOX, ELEMENT_SIZE = (128, 32, 32), 1.0 / 16.0
    WORKLOAD = ("configs[1]: SolidMechanics NeoHookean H1 p=2 hex, 32768 hexes / ~840k vdofs / 884736 "
                "qpts per GPU; step = residual with derivative store + one Jacobian action")
    STRONG_WORKLOAD = ("strong-scaling variant (SB_BENCH_SCALING=strong): one fixed 128x32x32-hex mesh split over "
                       "the GPUs; SolidMechanics NeoHookean H1 p=2 hex; step = residual with derivative store + "
                       "one Jacobian action")


def _split_even(n, parts):
    base, rem = divmod(n, parts)
    return [base + (1 if i < rem else 0) for i in range(parts)]


def build_structured_problem(device):
    """config B on one GPU: the structured cube, built directly (no partition, P = identity)"""
    import torch
    from serac_b200 import capi
    from serac_b200.structured import StructuredHexMesh
    p = ORDER
    n = STRONG_BOX if STRONG_SCALING else PER_GPU_BOX
    mesh = StructuredHexMesh(*n, *(n[d] * ELEMENT_SIZE for d in range(3)))
    tab, tab1 = mesh.element_table(p), mesh.element_table(1)
    G = capi.ElementRestriction(tab, 3, mesh.num_nodes(p))
    G_shape = capi.ElementRestriction(tab1, 3, mesh.num_nodes(1))
    nodes = torch.as_tensor(mesh.vertex_coordinates(), dtype=torch.float64, device=device)
    X_q, J_q = capi.geometric_factors(capi.CUBE, 1, p + 1, G_shape, nodes)
    integral = capi.Integral(f"solid_neohookean_hex_p{p}", NEOHOOKEAN, X_q, J_q, mesh.ne)
    F = capi.Functional(G, [G_shape, G, G])
    F.AddDomainIntegral(integral, [0, 1, 2])
    vs, ss = G.LSize(), G_shape.LSize()
    g = np.arange(vs, dtype=np.float64)
    host = dict(U=synthetic_u(g), A=np.zeros(vs), S=np.zeros(ss), dU=synthetic_du(g))
    return dict(capi=capi, G=G, G_shape=G_shape, integral=integral, F=F, host=host, ne=mesh.ne,
                nq=mesh.ne * (p + 1)**3, vdofs=vs, tdofs=vs, tdofs_shape=ss, p=p, keep=(nodes, X_q, J_q),
                mesh=f"structured {n[0]}x{n[1]}x{n[2]} hexes, one GPU")


def bind_to_gpu_numa_node(local_rank):
    """pin this process (and, by first touch, the host buffers it allocates next) to the NUMA node of its GPU"""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:  # noqa: BLE001  (no sysfs entry, no permission: stay where we are)
        pass
    return None


def copies_only_ms(dev, rank, world, ups, downs, reps=20):
    """the host -> device and device -> host copies of one step, up and down concurrently on two streams, nothing else:
    max over ranks of the time per step when one rank copies at a time (`one_rank_at_a_time`) and when all do
    (`all_ranks_at_once`) -- the floor the host links put under the end-to-end step"""
    import torch
    import torch.distributed as dist
    d_up = [torch.empty_like(h, device=dev) for h in ups]
    d_dn = [torch.zeros_like(h, device=dev) for h in downs]
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def once():
        with torch.cuda.stream(s_up):
            for d, h in zip(d_up, ups):
                d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s_dn):
            for d, h in zip(d_dn, downs):
                h.copy_(d, non_blocking=True)

    def timed():
        for _ in range(2):
            once()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            once()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3

    keep = [h.clone() for h in downs]  # the downloads above overwrite the step's results: put them back afterwards
    solo = 0.0
    for turn in range(world):
        dist.barrier()
        if turn == rank:
            solo = timed()
    dist.barrier()
    together = timed()
    for h, k in zip(downs, keep):
        h.copy_(k)
    t = torch.tensor([solo, together], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"one_rank_at_a_time": float(t[0].item()), "all_ranks_at_once": float(t[1].item())}


def synthetic_u(gid):
    """displacement as a function of the GLOBAL vdof id (shared dofs agree across ranks; small against the element size)"""
    return 0.016 * ELEMENT_SIZE * np.sin(0.37 * np.asarray(gid, dtype=np.float64) + 1.0)


def synthetic_du(gid):
    return 0.5 + 0.5 * np.sin(0.11 * np.asarray(gid, dtype=np.float64))


def global_parity(P, dev, rank, world, r_T, dr_T):
    """N > 1, untimed: rank 0 evaluates the SAME problem on the whole (unpartitioned) mesh, every rank compares its true
    dofs of the partitioned residual and Jacobian action with it.  -> dict for the JSON line (relative L2 over all
    dofs), or a reason when the global problem does not fit beside rank 0's partition."""
    import torch
    import torch.distributed as dist
    from serac_b200 import capi
    mesh, p = P["global_mesh"], P["p"]
    n = mesh.num_nodes(p) * 3
    need = mesh.ne * (p + 1)**3 * 360.0 * 1.6 + 64.0 * n
    free = torch.cuda.mem_get_info(dev)[0]
    fits = torch.tensor([1 if (rank != 0 or need < 0.8 * free) else 0], device=dev)
    dist.broadcast(fits, 0)
    if int(fits.item()) == 0:
        return {"skipped": f"the global problem ({mesh.ne} hexes, {need / 1e9:.0f} GB) does not fit beside rank 0's "
                           "partition; partitioned-vs-global parity at this order: tools/par_check.py"}
    r_g = torch.empty(n, dtype=torch.float64, device=dev)
    dr_g = torch.empty(n, dtype=torch.float64, device=dev)
    if rank == 0:
        t64 = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
        tab, tab1 = mesh.element_table(p), mesh.element_table(1)
        G = capi.ElementRestriction(tab, 3, mesh.num_nodes(p))
        G1 = capi.ElementRestriction(tab1, 3, mesh.num_nodes(1))
        nodes = t64(mesh.vertex_coordinates())
        X_q, J_q = capi.geometric_factors(capi.CUBE, 1, p + 1, G1, nodes)
        I = capi.Integral(f"solid_neohookean_hex_p{p}", NEOHOOKEAN, X_q, J_q, mesh.ne)
        F = capi.Functional(G, [G1, G, G])
        F.AddDomainIntegral(I, [0, 1, 2])
        gid = np.arange(n)
        F(0.0, [t64(np.zeros(mesh.num_nodes(1) * 3)), t64(synthetic_u(gid)), t64(np.zeros(n))], wrt=1, out=r_g)
        F.GradientMult(1, t64(synthetic_du(gid)), out=dr_g)
        torch.cuda.synchronize()
        del F, I, X_q, J_q, G, G1
    dist.broadcast(r_g, 0)
    dist.broadcast(dr_g, 0)
    gt = torch.as_tensor(P["part"].global_true_vdofs(), dtype=torch.int64, device=dev)
    acc = torch.stack([((r_T - r_g[gt])**2).sum(), (r_g[gt]**2).sum(), ((dr_T - dr_g[gt])**2).sum(),
                       (dr_g[gt]**2).sum()])
    dist.all_reduce(acc)
    e_r, e_dr = float((acc[0] / acc[1]).sqrt().item()), float((acc[2] / acc[3]).sqrt().item())
    return {"residual_rel_l2": e_r, "jvp_rel_l2": e_dr, "tolerance": 1e-12, "ok": bool(e_r < 1e-12 and e_dr < 1e-12),
            "reference": f"the same Functional on the whole {mesh.n[0]}x{mesh.n[1]}x{mesh.n[2]} mesh, one GPU (rank 0)"}


def build_partitioned_problem(device, rank, world):
    """N > 1: the config-A box per GPU (weak scaling), boxes laid out 2x1x1 / 2x2x1 / 2x2x2, one
    partition per rank with the shared-dof exchange P / P^T (mfem/fem/pfespace.cpp:3599-3663) on the box faces."""
    import torch
    from serac_b200 import capi
    from serac_b200.parallel import HaloExchange, ParFunctional
    from serac_b200.structured import Partition, StructuredHexMesh, factor_ranks
    p = ORDER
    ranks = factor_ranks(world)
    if os.environ.get("SB_BENCH_RANKS"):  # experiments: e.g. "1,2,1" cuts the two boxes across their large face
        ranks = tuple(int(x) for x in os.environ["SB_BENCH_RANKS"].split(","))
        assert int(np.prod(ranks)) == world
    n = tuple(PER_GPU_BOX[d] * ranks[d] for d in range(3))
    per_gpu = PER_GPU_BOX
    if STRONG_SCALING:  # SURVEY.md section 8d "also report": one fixed mesh split over the GPUs (near-equal boxes)
        n, per_gpu = STRONG_BOX, tuple(-(-STRONG_BOX[d] // ranks[d]) for d in range(3))
    mesh = StructuredHexMesh(*n, *(n[d] * ELEMENT_SIZE for d in range(3)))
    # owned-first local numbering: the T-vector is the prefix of the L-vector, so P / P^T run in place
    inplace = (os.environ.get("SB_HALO_TRANSPORT", "peer") == "peer" and os.environ.get("SB_HALO_OVERLAP", "1") != "0"
               and os.environ.get("SB_HALO_INPLACE", "1") != "0")
    part = Partition(mesh, ranks, rank, p, 3, owned_first=inplace)
    part1 = Partition(mesh, ranks, rank, 1, 3, owned_first=inplace)
    # elements all of whose nodes this rank owns come first: they are evaluated while the input exchange is in flight
    # measured: no gain at 2 GPUs, -2.5% at 8; its half-exchanges need the flag protocol (SB_HALO_LL=0)
    split = inplace and os.environ.get("SB_HALO_SPLIT", "0") != "0" and os.environ.get("SB_HALO_LL", "1") == "0"
    order, n_independent = part.halo_independent_first() if split else (None, 0)
    tab, tab1 = part.element_table(p, element_order=order), part1.element_table(1, element_order=order)
    G = capi.ElementRestriction(tab, 3, part.num_local_nodes)
    G_shape = capi.ElementRestriction(tab1, 3, part1.num_local_nodes)
    nodes = torch.as_tensor(part1.vertex_coordinates(), dtype=torch.float64, device=device)
    X_q, J_q = capi.geometric_factors(capi.CUBE, 1, p + 1, G_shape, nodes)
    integral = capi.Integral(f"solid_neohookean_hex_p{p}", NEOHOOKEAN, X_q, J_q, tab.shape[0])
    F = capi.Functional(G, [G_shape, G, G])
    F.AddDomainIntegral(integral, [0, 1, 2])
    # one halo object per field whose exchange can be in flight at the same time (u, a, the JVP direction, r, dr)
    lists, lists1 = part.halo_lists(), part1.halo_lists()
    halo1 = HaloExchange(lists1, device)
    halo, halo_a, halo_du, halo_r = (HaloExchange(lists, device) for _ in range(4))
    halos = [halo1, halo, halo_a, halo_du, halo_r]
    transport = "nccl"
    if os.environ.get("SB_HALO_TRANSPORT", "peer") == "peer":
        ok = [h.enable_peer() for h in halos]
        transport = "peer" if all(ok) else "nccl"
    overlap = transport == "peer" and os.environ.get("SB_HALO_OVERLAP", "1") != "0"
    if inplace and transport != "peer":
        raise RuntimeError("owned-first partitions were built for the peer transport, which is unavailable: rerun with "
                           "SB_HALO_INPLACE=0")
    if overlap:
        # puts before waits for the three inputs; P(dU) prefetched and P^T(r) on a side stream
        PF = ParFunctional(F, [halo1, halo, halo_a], halo, du_halos={1: halo_du}, out_halo=halo_r, inplace=inplace,
                           independent_elements=n_independent)
    else:
        PF = ParFunctional(F, [halo1, halo, halo], halo)
    # globally consistent synthetic inputs: a function of the global vdof id, so shared dofs agree across ranks
    gt, gt1 = part.global_true_vdofs().astype(np.float64), part1.global_true_vdofs().astype(np.float64)
    host = dict(U=synthetic_u(gt), A=np.zeros(len(gt)), S=np.zeros(len(gt1)), dU=synthetic_du(gt))
    return dict(capi=capi, G=G, G_shape=G_shape, integral=integral, F=F, PF=PF, halo=halo, halo1=halo1, halos=halos,
                part=part, part1=part1, global_mesh=mesh,
                overlap=overlap, inplace=inplace, n_independent=n_independent, host=host,
                ne=tab.shape[0], nq=tab.shape[0] * (p + 1)**3, vdofs=G.LSize(), tdofs=halo.tsize, tdofs_shape=halo1.tsize,
                global_vdofs=mesh.num_nodes(p) * 3, p=p, transport=transport, ranks=ranks, keep=(nodes, X_q, J_q),
                mesh=f"structured {n[0]}x{n[1]}x{n[2]} hexes, {ranks[0]}x{ranks[1]}x{ranks[2]} boxes of "
                     f"{per_gpu[0]}x{per_gpu[1]}x{per_gpu[2]}" + (" (strong scaling: fixed mesh)" if STRONG_SCALING else ""))


ASYNC_E2E = os.environ.get("SB_BENCH_E2E", "async") != "sync"


def run_product(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        P = build_partitioned_problem(dev, rank, world)
        call = P["PF"]
    else:
        P = build_problem(dev) if BENCH_CONFIG != "b" else build_structured_problem(dev)
        call = P["F"]
    capi, F, G, integral = P["capi"], P["F"], P["G"], P["integral"]
    t64 = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
    overlap = world > 1 and P.get("overlap", False)
    inplace = world > 1 and P.get("inplace", False)

    def tl(host, lsize):
        """in-place exchanges: an L-sized device buffer whose first tsize entries are the T-vector"""
        buf = torch.zeros(lsize, dtype=torch.float64, device=dev)
        buf[:len(host)] = t64(host)
        return buf

    if inplace:
        U, A, dU = (tl(P["host"][k], P["vdofs"]) for k in ("U", "A", "dU"))
        S = tl(P["host"]["S"], P["G_shape"].LSize())
    else:
        U, A, S, dU = (t64(P["host"][k]) for k in ("U", "A", "S", "dU"))  # T-vectors (= L-vectors at one rank)
    r = torch.empty(P["tdofs"], dtype=torch.float64, device=dev)
    dr = torch.empty_like(r)
    r_E = torch.empty(G.ESize(), dtype=torch.float64, device=dev)
    views = {}

    # experiment (off): the residual's scatter does not gate the Jacobian action (which needs the stored derivatives
    # only), so it can run on a second stream: sb_functional_evaluate_begin (all elements) / _end (the scatter).  Measured:
    # no gain (0.2796 vs 0.2788 ms, profiles/r3_experiments.md) -- the action kernel is persistent and holds every SM, so
    # the scatter's CTAs wait for it anyway
    fork_scatter = world == 1 and os.environ.get("SB_BENCH_FORK_SCATTER", "0") != "0"
    aux = torch.cuda.Stream() if fork_scatter else None
    ev_fork, ev_join = torch.cuda.Event(), torch.cuda.Event()

    def step_device():
        if fork_scatter:
            cur = torch.cuda.current_stream()
            call.evaluate_begin(0.0, [S, U, A], wrt=1, split_element=P["ne"])
            ev_fork.record(cur)
            with torch.cuda.stream(aux):
                aux.wait_event(ev_fork)
                call.evaluate_end(out=r)
                ev_join.record(aux)
            call.GradientMult(1, dU, out=dr)
            cur.wait_event(ev_join)
            return
        if overlap:
            call.prefetch_direction(1, dU)  # side stream: the exchange of the JVP direction runs next to the evaluation
        if inplace:  # results are the prefixes of the Functional's own L-vectors
            views["r"] = call(0.0, [S, U, A], wrt=1)
            views["dr"] = call.GradientMult(1, dU)
        else:
            call(0.0, [S, U, A], wrt=1, out=r)
            call.GradientMult(1, dU, out=dr)
        if overlap:
            call.join()

    # --- warm-up (also allocates the derivative buffer), then capture the step in a CUDA graph ---
    for _ in range(max(args.warmup, 3)):
        step_device()
    torch.cuda.synchronize()
    capi.lib().sb_launch_count_reset()
    step_device()
    torch.cuda.synchronize()
    launches_per_step = capi.lib().sb_launch_count()
    # the NCCL point-to-point transport is driven from the host, so only the peer-memory transport is graph-captured
    use_graph = not args.no_graph and (world == 1 or P["transport"] == "peer")
    if use_graph:
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            step_device()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            with torch.cuda.graph(graph, stream=side):
                step_device()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        if use_graph:
            graph.replay()
        else:
            step_device()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    ms_per_step = ms_total / args.steps
    if world > 1 and P["transport"] == "peer":
        assert all(h.peer_status() == 0 for h in P["halos"]), "a peer-memory flag wait timed out"

    # --- per-kernel durations inside a second timed loop (events on the launching stream; rank-local kernels only) ---
    UL, AL, SL, dUL = (U, A, S, dU) if world == 1 else (P["PF"]._L[1], P["PF"]._L[2], P["PF"]._L[0], P["PF"]._duL[1])
    rL = torch.empty(P["vdofs"], dtype=torch.float64, device=dev)
    fields = [P["G_shape"].field(SL), G.field(UL), G.field(AL)]
    du_field = G.field(dUL)
    psteps = min(args.steps, 200 if BENCH_CONFIG != "b" else 10)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(psteps)]
    for k in range(psteps):
        ev[k][0].record()
        # the same kernels, in the same E-vector layout ([element][node][comp]), as sb_functional_* launches in the step
        integral.Mult(0.0, fields, r_E, wrt=1, node_major=True)
        ev[k][1].record()
        G.Scatter(r_E, rL, node_major=True)
        ev[k][2].record()
        integral.GradientMult(1, du_field, r_E, node_major=True)
        ev[k][3].record()
        G.Scatter(r_E, rL, node_major=True)
        ev[k][4].record()
    torch.cuda.synchronize()
    phase_ms = [float(np.mean([ev[k][i].elapsed_time(ev[k][i + 1]) for k in range(psteps)])) for i in range(4)]
    halo_ms = None
    if world > 1:  # the exchange alone: P of the three inputs + P^T of the output, as in one evaluation
        dist.barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        for _ in range(psteps):
            if inplace:
                P["halo1"].P_inplace(S)
                P["halo"].P_inplace(U)
                P["halo"].P_inplace(A)
                P["halo"].Pt_inplace(rL)
            else:
                P["halo1"].P(S, SL)
                P["halo"].P(U, UL)
                P["halo"].P(A, AL)
                P["halo"].Pt(rL, r)
        h1.record()
        torch.cuda.synchronize()
        halo_ms = h0.elapsed_time(h1) / psteps
        dist.barrier()
        # restore the outputs the sanity check below compares against
        step_device()
        torch.cuda.synchronize()
    if inplace:
        r, dr = views["r"].clone(), views["dr"].clone()

    # --- end to end through the host-buffer API (pinned host memory in, host memory out) ---
    # every rank's host buffers live on the NUMA node its GPU hangs off (first touch after binding the process there):
    # eight ranks pulling from node 0 share one socket's memory controllers and one inter-socket link
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    pin = lambda a: torch.as_tensor(a, dtype=torch.float64).clone().pin_memory()
    hU, hA, hS, hdU = (pin(P["host"][k]) for k in ("U", "A", "S", "dU"))
    hr = torch.empty(P["tdofs"], dtype=torch.float64).pin_memory()
    hdr = torch.empty(P["tdofs"], dtype=torch.float64).pin_memory()

    if world == 1:
        # sb_functional_evaluate / sb_functional_jvp with host pointers: the copies happen inside the C ABI call
        nU, nA, nS, ndU, nr, ndr = (x.numpy() for x in (hU, hA, hS, hdU, hr, hdr))

        def step_host():
            # SB_HOST_ASYNC: r's device -> host copy overlaps the Jacobian action's host -> device copy and kernels
            F(0.0, [nS, nU, nA], wrt=1, out=nr, async_host=ASYNC_E2E)
            F.GradientMult(1, ndU, out=ndr, async_host=ASYNC_E2E)
            F.wait()
    else:
        if inplace:
            dS, dUd, dA, ddU = (torch.zeros_like(x) for x in (S, U, A, dU))  # L-sized, T-vector in the prefix
            ts, ts1 = P["tdofs"], P["tdofs_shape"]
            # the copies run on their own streams, as sb_functional_* does with SB_HOST_ASYNC at one rank: the direction
            # of the Jacobian action goes up during the evaluation, the residual comes down during the Jacobian action
            up, down = torch.cuda.Stream(), torch.cuda.Stream()
            ev = {k: torch.cuda.Event() for k in ("in", "du", "r", "dr", "consumed", "copied")}

            # the device work between the copies is two CUDA graphs (evaluation with its exchanges; Jacobian action with
            # its exchanges), so the host side of a step is a dozen calls -- at one rank the C ABI does the same inside
            # sb_functional_evaluate / sb_functional_jvp
            host_graphs = None
            if use_graph:
                gA, gB, out_views = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), {}
                cap = torch.cuda.Stream()
                with torch.cuda.stream(cap):
                    for _ in range(2):  # warm-up on the capture stream
                        call(0.0, [dS, dUd, dA], wrt=1)
                        call.GradientMult(1, ddU)
                        call.join()
                    torch.cuda.synchronize()
                    dist.barrier()
                    with torch.cuda.graph(gA, stream=cap):
                        out_views["r"] = call(0.0, [dS, dUd, dA], wrt=1)
                        call.join()
                    with torch.cuda.graph(gB, stream=cap):
                        out_views["dr"] = call.GradientMult(1, ddU)
                torch.cuda.synchronize()
                dist.barrier()
                host_graphs = (gA, gB)

            def step_host():
                main = torch.cuda.current_stream()
                up.wait_event(ev["consumed"])  # the previous step's kernels are done with the staging buffers
                with torch.cuda.stream(up):
                    dS[:ts1].copy_(hS, non_blocking=True)
                    dUd[:ts].copy_(hU, non_blocking=True)
                    dA[:ts].copy_(hA, non_blocking=True)
                    ev["in"].record(up)
                    ddU[:ts].copy_(hdU, non_blocking=True)
                    ev["du"].record(up)
                main.wait_event(ev["in"])
                main.wait_event(ev["copied"])  # the previous step's downloads have left the Functional's L-vectors
                if host_graphs:
                    host_graphs[0].replay()
                    rv = out_views["r"]
                else:
                    rv = call(0.0, [dS, dUd, dA], wrt=1)
                    call.join()  # the residual's P^T ran on the side stream
                ev["r"].record(main)
                main.wait_event(ev["du"])
                if host_graphs:
                    host_graphs[1].replay()
                    drv = out_views["dr"]
                else:
                    drv = call.GradientMult(1, ddU)
                ev["dr"].record(main)
                ev["consumed"].record(main)
                with torch.cuda.stream(down):
                    down.wait_event(ev["r"])
                    hr.copy_(rv, non_blocking=True)
                    down.wait_event(ev["dr"])
                    hdr.copy_(drv, non_blocking=True)
                    ev["copied"].record(down)
                down.synchronize()
        else:
            dS, dUd, dA, ddU = (torch.empty_like(x, device=dev) for x in (hS, hU, hA, hdU))
            d_r, d_dr = torch.empty_like(r), torch.empty_like(dr)

            def step_host():
                dS.copy_(hS, non_blocking=True)
                dUd.copy_(hU, non_blocking=True)
                dA.copy_(hA, non_blocking=True)
                call(0.0, [dS, dUd, dA], wrt=1, out=d_r)
                hr.copy_(d_r, non_blocking=True)
                ddU.copy_(hdU, non_blocking=True)
                call.GradientMult(1, ddU, out=d_dr)
                if overlap:
                    call.join()
                hdr.copy_(d_dr, non_blocking=True)
                torch.cuda.current_stream().synchronize()

    for _ in range(3):
        step_host()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # long enough that one nvidia-smi poll of the clock sampler averages out (config B: a step moves GBs over PCIe)
    e2e_steps = max(3, min(args.steps, 200 if BENCH_CONFIG != "b" else 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    clocks = sampler.stop() if rank == 0 else None

    # N > 1: what the host links allow -- the step's copies and nothing else, by this rank alone and by all ranks at once
    # (untimed part of the run; the box's aggregate host <-> device bandwidth, not the kernels, bounds e2e at 8 GPUs:
    # profiles/r2x_pcie_probe_n8.json)
    copy_ceiling = None
    if world > 1 and BENCH_CONFIG != "b":
        copy_ceiling = copies_only_ms(dev, rank, world, (hS, hU, hA, hdU), (hr, hdr))

    # sanity: host path and device path agree bit for bit (same kernels, same data)
    if world > 1 and P["transport"] == "peer":
        assert all(h.peer_status() == 0 for h in P["halos"]), "a wait for a neighbour timed out in the host-buffer steps"
    # (compared on every rank; a mismatch is re-run once, reported in the JSON line either way, and fatal if it repeats)
    same_r, same_dr = np.array_equal(hr.numpy(), r.cpu().numpy()), np.array_equal(hdr.numpy(), dr.cpu().numpy())
    host_check = {"equal": bool(same_r and same_dr), "attempts": 1}
    flag = torch.tensor([0.0 if (same_r and same_dr) else 1.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    if flag.item() > 0:  # some rank saw a difference: everybody repeats both paths (the exchanges are collective)
        first = dict(rank=rank, r_equal=bool(same_r), dr_equal=bool(same_dr),
                     dr_entries_off=int((hdr.numpy() != dr.cpu().numpy()).sum()),
                     dr_max_abs_diff=float(np.abs(hdr.numpy() - dr.cpu().numpy()).max()))
        step_device()
        torch.cuda.synchronize()
        dev2_r = (views["r"] if inplace else r)[:hr.numel()].cpu().numpy().copy()
        dev2_dr = (views["dr"] if inplace else dr)[:hdr.numel()].cpu().numpy().copy()
        step_host()
        torch.cuda.synchronize()
        again = bool(np.array_equal(hr.numpy(), dev2_r) and np.array_equal(hdr.numpy(), dev2_dr))
        device_repeatable = bool(np.array_equal(dev2_dr, dr.cpu().numpy()[:hdr.numel()]))
        firsts = [None] * world
        if world > 1:
            dist.all_gather_object(firsts, first)
        else:
            firsts = [first]
        host_check = {"equal": again, "attempts": 2, "device_path_repeatable": device_repeatable,
                      "first_attempt": [f for f in firsts if not (f["r_equal"] and f["dr_equal"])]}
        ok = torch.tensor([1.0 if again else 0.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() < 1:
            raise AssertionError(f"rank {rank}: host-buffer path differs from the device path twice in a row: {host_check}")
    assert np.isfinite(hr.numpy()).all() and np.abs(hdr.numpy()).max() > 0

    sums = torch.tensor([P["tdofs"], P["tdofs_shape"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(sums)
    vdofs_total, shape_total = int(sums[0].item()), int(sums[1].item())
    parity = None
    if world > 1:
        assert vdofs_total == P["global_vdofs"], "true dofs of the partitions do not add up to the global space"
        parity = global_parity(P, dev, rank, world, r[:P["tdofs"]], dr[:P["tdofs"]])
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    peak, peak_src = measured_peaks()
    # algorithmic bytes of the Jacobian action per quadrature point: du + dr L-vectors (8*3*p^3/(p+1)^3
    # each), 81 stored derivatives (648 B), int32 restriction table (4 B)  = 666 B/qpt at p=2
    p = P["p"]
    V = 8.0 * 3 * p**3 / (p + 1)**3
    # the stored derivative: 81 doubles per point in the reference's layout (648 B); the NeoHookean tangent is symmetric
    # and this library stores its upper triangle, 45 doubles (360 B).  The roofline is taken against the
    # bytes THIS algorithm has to move; the same time against the reference's 648 B is reported beside it.
    stored = P["integral"].derivative_stored_entries(1)
    full = P["integral"].derivative_entries(1)
    D, D_ref = 8.0 * stored, 8.0 * full
    jvp_bytes = (2 * V + D + 4.0) * P["nq"]
    jvp_ms = phase_ms[2]
    achieved = jvp_bytes / (jvp_ms * 1e-3) / 1e9
    eval_bytes = (3 * V + 0.9 + 96.0 + D + 4.0) * P["nq"]  # residual + derivative store, ~770 B/qpt at 81 entries (§8d)
    ref_layout = {"derivative_doubles_per_qpt": {"stored": stored, "reference_layout": full},
                  "jvp_GBps_if_reference_layout_bytes": (2 * V + D_ref + 4.0) * P["nq"] / (jvp_ms * 1e-3) / 1e9,
                  "evaluate_GBps_if_reference_layout_bytes":
                      (3 * V + 0.9 + 96.0 + D_ref + 4.0) * P["nq"] / (phase_ms[0] * 1e-3) / 1e9}
    out = {
        "metric": METRIC, "value": vdofs_total / (ms_per_step * 1e-3) / 1e9, "unit": "GDOF/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if STRONG_SCALING else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": STRONG_WORKLOAD if STRONG_SCALING else WORKLOAD,
                   "mesh": P["mesh"], "vdofs_total": vdofs_total, "vdofs_per_gpu": P["vdofs"],
                   "qpts_per_gpu": P["nq"],
                   "l2_policy": f"inputs larger than L2 ({D * P['nq'] / 1e6:.0f} MB of stored derivatives streamed each "
                                "step)",
                   "step_launch": ("CUDA graph replay" if use_graph else "kernel by kernel")
                                  + ("; the residual's scatter runs on a second stream under the Jacobian-action kernel"
                                     if fork_scatter else ""),
                   "parallelism": f"1 mesh partition per GPU x{world}"
                                  + (f", shared-dof exchange over {P['transport']}"
                                     + (", in place (owned-first numbering), overlapped on a side stream"
                                        if P.get("inplace") else "")
                                     + (f", {P.get('n_independent', 0)} halo-independent elements evaluated during the "
                                        "input exchange" if P.get("n_independent") else "") if world > 1 else "")},
        "e2e": {"value": vdofs_total / e2e_s / 1e9, "unit": "GDOF/s",
                "h2d_bytes_per_step": int(8 * (3 * vdofs_total + shape_total)),
                "d2h_bytes_per_step": int(8 * 2 * vdofs_total), "ms_per_step": e2e_s * 1e3,
                "host_buffers": "pinned" + (f", bound to the GPU's NUMA node (rank 0: node {numa})" if numa is not None
                                            else ""),
                **({"copies_only_ms": copy_ceiling} if copy_ceiling else {})},
        "gpu_launches": int(launches_per_step * args.steps),
        "phases_ms": {"evaluate_with_derivative_store": phase_ms[0], "scatter_r": phase_ms[1], "jvp": phase_ms[2],
                      "scatter_dr": phase_ms[3]},
        "qpts_per_s": P["nq"] * world / (ms_per_step * 1e-3),
        "roofline": {"kernel": f"jvp_stream_kernel<3,{p + 1},{p + 1},3,{p + 1},3,{P['integral'].derivative_mask(1)},*>",
                     "bound": "hbm", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": jvp_bytes,
                     "traffic": JVP_NCU_TRAFFIC_BYTES if BENCH_CONFIG != "b" else None,
                     "traffic_source": NCU_TRAFFIC_SOURCE if BENCH_CONFIG != "b" else None,
                     "limiter": JVP_LIMITER if BENCH_CONFIG != "b" else None},
        "roofline_evaluate": {"kernel": "evaluation_kernel<wrt=1>", "bound": "hbm",
                              "achieved": eval_bytes / (phase_ms[0] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                              "frac": eval_bytes / (phase_ms[0] * 1e-3) / 1e9 / peak,
                              "algorithmic_bytes_per_launch": eval_bytes,
                              "traffic": EVAL_NCU_TRAFFIC_BYTES if BENCH_CONFIG != "b" else None,
                              "traffic_source": NCU_TRAFFIC_SOURCE if BENCH_CONFIG != "b" else None,
                              "limiter": EVAL_LIMITER if BENCH_CONFIG != "b" else None},
        "derivative_storage": ref_layout,
        "clocks": clocks,
        "host_path_equals_device_path": host_check,
    }
    if halo_ms is not None:
        out["phases_ms"]["halo_exchange_3P_1Pt"] = halo_ms
    if parity is not None:
        out["parity_vs_global"] = parity
    if args.gpus == 1 and world == 1 and not args.no_solver_layer and not args.no_graph:  # (--no-graph = ncu captures)
        # outside the timed step: the Newton step's linear algebra on the same problem.  Reported
        # beside the headline, never part of it; a failure here is recorded, not raised
        try:
            out["solver_layer"] = solver_layer_sample(P, dev, peak)
        except Exception as e:  # noqa: BLE001
            out["solver_layer"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if args.gpus == 1 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_sample()
    print(json.dumps(out))


def solver_layer_sample(P, dev, peak):
    """SpMV / CG / GMRES on the eliminated config-A Jacobian (include/serac_b200_solvers.h): CUDA-event times on the
    stream the kernels are launched on, algorithmic bytes = 12 B per nonzero + 16 B per row for y = A x."""
    import torch
    capi, F, G = P["capi"], P["F"], P["G"]
    t64 = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
    U, A, S, dU = (t64(P["host"][k]) for k in ("U", "A", "S", "dU"))
    n = G.LSize()
    F(0.0, [S, U, A], wrt=1)
    pat, vals = F.assemble(1)

    def timed(fn, reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # assemble(Gradient&) of BASELINE configs[2] (every Newton step pays it): the pattern and the plan exist after the
    # first call, so this is the fused element-stiffness -> CSR kernel alone (tools/assemble_timing.py
    # has the comparison with the element-matrix route)
    timed(lambda: F.assemble(1), 2)
    assemble_ms = timed(lambda: F.assemble(1), 5)
    pat, vals = F.assemble(1)
    ess = np.arange(0, 3 * (n // 3 // 64), dtype=np.int32)
    capi.eliminate_rows_cols(pat, vals, ess, want_Ae=False)
    dinv = capi.jacobi_setup(capi.csr_diagonal(pat, vals))
    y = torch.empty(n, dtype=torch.float64, device=dev)

    timed(lambda: capi.csr_spmv(pat, vals, dU, out=y), 3)
    spmv_ms = timed(lambda: capi.csr_spmv(pat, vals, dU, out=y), 10)
    nnz = int(pat.nnz)
    spmv_gbs = (12.0 * nnz + 16.0 * n) / (spmv_ms * 1e-3) / 1e9
    b = y.clone()
    res = {"workload": "eliminated NeoHookean Jacobian of config A", "rows": int(n), "nnz": nnz,
           "assemble_ms": assemble_ms,
           "assemble_roofline": {"kernel": "assemble_rows_kernel", "bound": "fp64", "flops": 14189395968.0,
                                 "achieved_tflops": 14189395968.0 / (assemble_ms * 1e-3) / 1e12, "peak_tflops": 37.1,
                                 "peak_source": "measured DFMA rate, profiles/r2a_fp64_peak_b200.json",
                                 "frac": 14189395968.0 / (assemble_ms * 1e-3) / 1e12 / 37.1,
                                 "limiter": "shared-memory pipe 55%, 15% of it bank conflicts of the row accumulation "
                                            "(profiles/r4_summary.md)"},
           "spmv_ms": spmv_ms, "spmv_gbs": spmv_gbs, "spmv_frac_of_hbm_peak": spmv_gbs / peak,
           "note": "outside the timed step; 50 iterations each, tolerances 0 so that exactly 50 run"}
    iters = 50
    # the same Jacobi-preconditioned CG matrix-free: [J_ii 0; 0 I] through Gradient::Mult (the streaming Jacobian action
    # + scatter) instead of the assembled matrix -- what a Newton-Krylov step costs when nothing is assembled
    op = capi.GradientOperator(F, 1, ess)
    for name, solver, operator in (
            ("cg_jacobi", capi.CGSolver(n, rel_tol=0.0, abs_tol=0.0, max_iter=iters, check_every=10), pat),
            ("cg_jacobi_matrix_free", capi.CGSolver(n, rel_tol=0.0, abs_tol=0.0, max_iter=iters, check_every=10), op),
            ("gmres50_jacobi", capi.GMRESSolver(n, kdim=50, rel_tol=0.0, abs_tol=0.0, max_iter=iters), pat)):
        kw = dict(values=vals) if operator is pat else {}
        solver.Mult(operator, b, dinv=dinv, **kw)  # warm-up
        capi.lib().sb_launch_count_reset()
        box = {}
        ms = timed(lambda: box.update(info=solver.Mult(operator, b, dinv=dinv, **kw)[1]), 1)
        info = box["info"]
        res[name] = {"iterations": int(info["final_iter"]), "ms_per_iteration": ms / max(int(info["final_iter"]), 1),
                     "launches": int(capi.lib().sb_launch_count()),
                     "residual_reduction": float(info["final_norm"] / info["initial_norm"])}
    return res


# ---------------------------------------------------------------------------------------------------------------------
# reference CPU arm (oracle/_ref = the reference's own kernels; the only place bench.py executes oracle/)
# ---------------------------------------------------------------------------------------------------------------------
def _reference_slab(job):
    """time `steps` reference steps (residual with derivative store + Jacobian action, gather/scatter included) on a
    slab of nx x 16 x 16 hexes of the config-A beam (element size 1/16)."""
    nx, steps, warmup = job
    from oracle import ref
    mesh = ref.Mesh.cartesian3d(nx, 16, 16, nx / 16.0, 1.0, 1.0)
    space, shape_space = ref.Space(mesh, 2, 3), ref.Space(mesh, 1, 3)
    I = ref.Integral("solid_neohookean_3d_p2", mesh, NEOHOOKEAN)
    rng = np.random.default_rng(42)
    U, dU = 1e-3 * rng.uniform(-1, 1, space.vsize), rng.uniform(0, 1, space.vsize)
    A, S = np.zeros(space.vsize), np.zeros(shape_space.vsize)

    def step():
        ins = [shape_space.gather(S), space.gather(U), space.gather(A)]
        r = space.scatter_add(I.mult(0.0, ins, space.esize, wrt=1), np.zeros(space.vsize))
        dr = space.scatter_add(I.jvp(1, space.gather(dU), space.esize), np.zeros(space.vsize))
        return r, dr

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return (time.perf_counter() - t0) / steps, space.vsize


def reference_run(cores, nx_per_core, steps, warmup):
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(_reference_slab, [(nx_per_core, steps, warmup)] * cores)
    t = max(r[0] for r in res)
    vdofs = sum(r[1] for r in res)
    return vdofs / t / 1e9, t, vdofs


def reference_cores():
    """processes of the reference arm: the largest divisor of the beam's 128 hexes along x that the host cores allow, so
    that the slabs tile config A's 128 x 16 x 16 beam exactly (one 1/R slab of the mesh per core)"""
    cores = max(1, min(os.cpu_count() or 1, 128))
    return max(d for d in (1, 2, 4, 8, 16, 32, 64, 128) if d <= cores)


CONFIG_A_VDOFS = 839619  # true dofs of the beam (the slabs' duplicated interface nodes are the same dofs)


def cpu_baseline_sample():
    """bounded sample for the product arm's line: one timed step of the whole config-A beam, one slab per core"""
    from oracle import ref
    if not ref.available():
        return {"value": None, "unit": "GDOF/s", "cores": 0, "kind": "reference",
                "sample": "oracle/_ref not built on this box"}
    cores = reference_cores()
    nx = 128 // cores
    _, t, vdofs = reference_run(cores, nx, 1, 1)
    return {"value": CONFIG_A_VDOFS / t / 1e9, "unit": "GDOF/s", "cores": cores, "kind": "reference",
            "sample": f"config A itself: {cores} processes x ({nx}x16x16 hexes = 1/{cores} of the 128x16x16 beam, "
                      f"{vdofs // cores} local vdofs each), 1 warm-up + 1 timed step of the reference's own kernels "
                      f"(oracle/_ref, -O3), {t:.2f} s/step; gather/scatter included, no halo exchange"}


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import ref
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libserac_ref.so not built"}))
        return
    cores = reference_cores()
    nx = 128 // cores
    steps = max(1, min(args.steps, 3))
    _, t, vdofs = reference_run(cores, nx, steps, 1)
    value = CONFIG_A_VDOFS / t / 1e9
    sample = (f"config A itself: {cores} processes x ({nx}x16x16 hexes = 1/{cores} of the 128x16x16 beam, "
              f"{vdofs // cores} local vdofs each), 1 warm-up + {steps} timed steps, {t:.2f} s/step; gather/scatter "
              "included, no halo exchange (flatters the CPU slightly)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GDOF/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": 1, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "mesh": "the 128x16x16-hex beam of config A cut into one slab per host core",
                   "vdofs_total": CONFIG_A_VDOFS, "vdofs_per_gpu": CONFIG_A_VDOFS, "qpts_per_gpu": 884736,
                   "parallelism": f"{cores} host processes, one mesh slab each"},
        "cpu_baseline": {"value": value, "unit": "GDOF/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-solver-layer", action="store_true",
                    help="skip the out-of-step sample of the Newton-step linear algebra (solver_layer key)")
    ap.add_argument("--no-graph", action="store_true", help="launch the step kernel by kernel (for ncu captures)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
