"""Host <-> device copy bandwidth of every rank, ALONE and with ALL RANKS COPYING AT ONCE -- the platform ceiling of the
bench's end-to-end number at N > 1 (every rank moves 20.3 MB up and 13.0 MB down per step at config A).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29540 \
        tools/pcie_probe_ranks.py

Per rank: pinned buffers after binding the process to the GPU's NUMA node (what bench.py does), copies of the bench's
sizes, timed with CUDA events; `solo` = the other ranks wait at a barrier, `all` = everybody copies in the same window.
Rank 0 prints one JSON object."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import bench

UP_BYTES, DOWN_BYTES, REPS = 20_300_000, 13_000_000, 30


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    numa = bench.bind_to_gpu_numa_node(local)
    h_up = torch.empty(UP_BYTES // 8, dtype=torch.float64).pin_memory()
    h_dn = torch.empty(DOWN_BYTES // 8, dtype=torch.float64).pin_memory()
    h_up.fill_(1.0)
    d_up, d_dn = torch.empty_like(h_up, device=dev), torch.ones(DOWN_BYTES // 8, dtype=torch.float64, device=dev)
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def up():
        with torch.cuda.stream(s_up):
            d_up.copy_(h_up, non_blocking=True)

    def down():
        with torch.cuda.stream(s_dn):
            h_dn.copy_(d_dn, non_blocking=True)

    def both():
        up()
        down()

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s_up.wait_event(e0)
        s_dn.wait_event(e0)
        for _ in range(REPS):
            fn()
        torch.cuda.current_stream().wait_stream(s_up)
        torch.cuda.current_stream().wait_stream(s_dn)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / REPS  # ms per repetition

    def barrier():
        if world > 1:
            dist.barrier()

    res = {}
    for name, fn, nbytes in (("h2d", up, UP_BYTES), ("d2h", down, DOWN_BYTES), ("both", both, UP_BYTES + DOWN_BYTES)):
        solo = 0.0
        for turn in range(world):  # one rank at a time
            barrier()
            if turn == rank:
                solo = timed(fn)
            barrier()
        barrier()
        together = timed(fn)  # everybody at once
        barrier()
        res[name] = dict(solo_GBps=nbytes / solo / 1e6, all_GBps=nbytes / together / 1e6, solo_ms=solo, all_ms=together)
    row = dict(rank=rank, numa=numa, **{f"{k}_{m}": round(v[m], 3) for k, v in res.items() for m in v})
    rows = [None] * world
    if world > 1:
        dist.all_gather_object(rows, row)
    else:
        rows = [row]
    if rank == 0:
        agg = {k: round(sum(r[f"{k}_all_GBps"] for r in rows), 1) for k in ("h2d", "d2h", "both")}
        print(json.dumps(dict(world=world, up_bytes=UP_BYTES, down_bytes=DOWN_BYTES, aggregate_all_GBps=agg, ranks=rows)))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
