"""Kernel durations of VQVAE.quantize (fused projections vs Conv2d projections), CUPTI via torch.profiler."""
import os, sys, copy, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from torch.profiler import ProfilerActivity, profile
from pero_pretraining_b200 import VQVAE
dev = torch.device("cuda:0")
class P(torch.nn.Module):
    def __init__(s, c): super().__init__(); s.out_channels = s.base_channels = c
    def forward(s, x): return x
torch.manual_seed(0)
C, K, D = 256, 8192, 256
a = VQVAE(P(C), P(C), K, D, 0.25, 0.99).to(dev).eval(); b = copy.deepcopy(a); b.fuse_projections = False
x = torch.randn(64, C, 1, 128, device=dev)
for name, m in (("fused", a), ("conv2d", b)):
    with torch.no_grad():
        for _ in range(3): m.quantize(x)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            m.quantize(x); torch.cuda.synchronize()
    evs = sorted([e for e in prof.events() if "cuda" in str(getattr(e, "device_type", "")).lower()], key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    print("==", name)
    for e in evs:
        print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:8.1f} {e.name[:100]}")
