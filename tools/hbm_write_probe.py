"""HBM bandwidth of streams with different read : write mixes on this GPU (CUDA events, buffers >> L2): the copy figure
in MEASURED_PEAKS.json is a 50 : 50 mix; the evaluation kernel with derivative store writes 84% of its bytes."""
import json
import torch

dev = torch.device("cuda:0")
N = 1 << 29  # doubles: 4 GiB per buffer
a = torch.empty(N, dtype=torch.float64, device=dev)
b = torch.empty(N, dtype=torch.float64, device=dev)
a.fill_(1.0)
b.fill_(2.0)


def timed(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


B = N * 8
out = {}
out["write_only_fill_GBps"] = B / timed(lambda: a.fill_(3.0)) / 1e9
out["write_only_memset_GBps"] = B / timed(lambda: a.zero_()) / 1e9
out["copy_read_plus_write_GBps"] = 2 * B / timed(lambda: b.copy_(a)) / 1e9
out["read_only_sum_GBps"] = B / timed(lambda: a.sum()) / 1e9
# 1 read : 5.25 writes (the evaluation kernel's 104 MB : 536 MB): a[:n] read, b written whole
n = int(N / 5.25)
src = a[:n]


def mix():
    b.zero_()
    src.sum()


out["eval_like_mix_sequential_GBps"] = (B + n * 8) / timed(mix) / 1e9
print(json.dumps(out))
