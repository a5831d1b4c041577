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
                nq=int(d["ne"]) * (p + 1)**dim, vdofs=vs, p=p, keep=(nodes, X_q, J_q))


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
    P = build_problem(dev)
    capi, F, G, integral = P["capi"], P["F"], P["G"], P["integral"]
    t64 = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
    U, A, S, dU = (t64(P["host"][k]) for k in ("U", "A", "S", "dU"))
    r = torch.empty(P["vdofs"], dtype=torch.float64, device=dev)
    dr = torch.empty_like(r)
    r_E = torch.empty(G.ESize(), dtype=torch.float64, device=dev)

    def step_device():
        F(0.0, [S, U, A], wrt=1, out=r)
        F.GradientMult(1, dU, out=dr)

    # --- warm-up (also allocates the derivative buffer), then capture the step in a CUDA graph ---
    for _ in range(max(args.warmup, 3)):
        step_device()
    torch.cuda.synchronize()
    capi.lib().sb_launch_count_reset()
    step_device()
    torch.cuda.synchronize()
    launches_per_step = capi.lib().sb_launch_count()
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        step_device()
        torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=side):
            step_device()
    torch.cuda.synchronize()
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
        if args.no_graph:
            step_device()
        else:
            graph.replay()
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

    # --- per-kernel durations inside a second timed loop (events on the launching stream) ---
    fields = [P["G_shape"].field(S), G.field(U), G.field(A)]
    du_field = G.field(dU)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]
    for k in range(args.steps):
        ev[k][0].record()
        integral.Mult(0.0, fields, r_E, wrt=1)
        ev[k][1].record()
        G.Scatter(r_E, r)
        ev[k][2].record()
        integral.GradientMult(1, du_field, r_E)
        ev[k][3].record()
        G.Scatter(r_E, dr)
        ev[k][4].record()
    torch.cuda.synchronize()
    phase_ms = [float(np.mean([ev[k][i].elapsed_time(ev[k][i + 1]) for k in range(args.steps)])) for i in range(4)]

    # --- end to end through the host-buffer C ABI (pinned host memory in, host memory out) ---
    pin = lambda a: torch.as_tensor(a, dtype=torch.float64).pin_memory().numpy()
    hU, hA, hS, hdU = (pin(P["host"][k]) for k in ("U", "A", "S", "dU"))
    hr = torch.empty(P["vdofs"], dtype=torch.float64).pin_memory().numpy()
    hdr = torch.empty(P["vdofs"], dtype=torch.float64).pin_memory().numpy()

    def step_host():
        F(0.0, [hS, hU, hA], wrt=1, out=hr)
        F.GradientMult(1, hdU, out=hdr)

    for _ in range(3):
        step_host()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_steps = max(3, min(args.steps, 20))
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

    # sanity: host path and device path agree bit for bit (same kernels, same data)
    assert np.array_equal(hr, r.cpu().numpy()) and np.array_equal(hdr, dr.cpu().numpy())

    if rank != 0:
        return
    peak, peak_src = measured_peaks()
    vdofs_total = P["vdofs"] * world
    # algorithmic bytes of the Jacobian action per quadrature point: du + dr L-vectors (8*3*p^3/(p+1)^3
    # each), 81 stored derivatives (648 B), int32 restriction table (4 B)  = 666 B/qpt at p=2
    p = P["p"]
    V = 8.0 * 3 * p**3 / (p + 1)**3
    jvp_bytes = (2 * V + 648.0 + 4.0) * P["nq"]
    jvp_ms = phase_ms[2]
    achieved = jvp_bytes / (jvp_ms * 1e-3) / 1e9
    out = {
        "metric": METRIC, "value": vdofs_total / (ms_per_step * 1e-3) / 1e9, "unit": "GDOF/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[1]: SolidMechanics NeoHookean H1 p=2 hex, beam-hex.mesh refined 4x "
                               "(32768 hexes, 839619 vdofs, 884736 qpts) per GPU; step = residual with derivative store "
                               "+ one Jacobian action", "vdofs_per_gpu": P["vdofs"], "qpts_per_gpu": P["nq"],
                   "l2_policy": "inputs larger than L2 (573 MB stored derivatives streamed each step)",
                   "step_launch": "kernel by kernel" if args.no_graph else "CUDA graph replay", "parallelism": f"1 partition per GPU x{world}"},
        "e2e": {"value": vdofs_total / e2e_s / 1e9, "unit": "GDOF/s",
                "h2d_bytes_per_step": int(8 * (3 * P["vdofs"] + P["G_shape"].LSize())),
                "d2h_bytes_per_step": int(8 * 2 * P["vdofs"]), "ms_per_step": e2e_s * 1e3},
        "gpu_launches": int(launches_per_step * args.steps),
        "phases_ms": {"evaluate_with_derivative_store": phase_ms[0], "scatter_r": phase_ms[1], "jvp": phase_ms[2],
                      "scatter_dr": phase_ms[3]},
        "roofline": {"kernel": "jvp_kernel<3,3,3,3,3,3,8>", "bound": "hbm", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": jvp_bytes, "traffic": None},
        "clocks": clocks,
    }
    if args.gpus == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_sample()
    print(json.dumps(out))


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


def cpu_baseline_sample():
    from oracle import ref
    if not ref.available():
        return {"value": None, "unit": "GDOF/s", "cores": 0, "kind": "reference",
                "sample": "oracle/_ref not built on this box"}
    cores = max(1, min(os.cpu_count() or 1, 128))
    nx = 2  # 2 x 16 x 16 = 512 hexes per core: ~6 s of reference AD evaluation per step per core
    value, t, vdofs = reference_run(cores, nx, 1, 1)
    return {"value": value, "unit": "GDOF/s", "cores": cores, "kind": "reference",
            "sample": f"{cores} processes x (2x16x16 hexes of the config-A beam, {vdofs // cores} vdofs each), 1 warm-up "
                      f"+ 1 timed step of the reference's own kernels (oracle/_ref), {t:.2f} s/step; no halo exchange"}


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import ref
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libserac_ref.so not built"}))
        return
    cores = max(1, min(os.cpu_count() or 1, 128))
    nx = 1
    steps = max(1, min(args.steps, 3))
    value, t, vdofs = reference_run(cores, nx, steps, 1)
    sample = (f"{cores} processes x (1x16x16 hexes of the config-A beam = {vdofs // cores} vdofs each), "
              f"1 warm-up + {steps} timed steps, {t:.2f} s/step; gather/scatter included, no halo exchange")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GDOF/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": 1, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[1] slabs: SolidMechanics NeoHookean H1 p=2 hex, reference CPU kernels, one "
                               "mesh slab per host core", "vdofs_total": vdofs},
        "cpu_baseline": {"value": value, "unit": "GDOF/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step kernel by kernel (for ncu captures)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
