"""Masked-CE forward + backward at the bench shape, alone on the GPU: recompute route vs kept-numerators route
(device time per call pair by CUDA events; L2 flushed or warm), and the kernels of each route from CUPTI."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from pero_pretraining_b200 import ops
dev = torch.device("cuda:0")
N, Dh, V = 8192, 512, 8192
h = torch.randn(N, Dh, device=dev); W = torch.randn(V, Dh, device=dev) * 0.04; b = torch.zeros(V, device=dev)
labels = torch.randint(0, V, (N,), device=dev)
rows = torch.from_numpy(np.flatnonzero(np.random.default_rng(0).random(N) < 0.15).astype(np.int32)).to(dev)
head = ops.PreparedHead(V, Dh, dev).prepare(W, b)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
M = rows.numel()

def run(keep):
    loss_sum, lse, ws = ops.masked_ce_fwd(h, rows, labels, head, keep_logits=keep)
    return ops.masked_ce_bwd(h, rows, labels, head, lse, None, 1.0 / M, ws=ws, ws_from_fwd=True, logits_in_ws=keep)

for keep in (False, True):
    for fl in (True, False):
        ts = []
        for i in range(25):
            if fl: flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(keep); e1.record(); torch.cuda.synchronize()
            if i >= 5: ts.append(e0.elapsed_time(e1) * 1e3)
        print(f"keep={keep} flush={fl}: {np.mean(ts):.1f} us (min {np.min(ts):.1f})")
from torch.profiler import ProfilerActivity, profile
for keep in (False, True):
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3): run(keep)
        torch.cuda.synchronize()
    evs = sorted([e for e in prof.events() if "cuda" in str(getattr(e, "device_type", "")).lower()], key=lambda e: e.time_range.start)
    n = len(evs) // 3
    t0 = evs[2 * n].time_range.start
    print(f"--- keep={keep}")
    for e in evs[2 * n:]:
        print(f"{e.time_range.start - t0:8.1f} {e.time_range.end - e.time_range.start:7.1f} {e.name[:90]}")
