"""Per-CTA wall-clock profile + worker-0 unit timeline of the four masked-CE GEMM launches at the bench shape."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from pero_pretraining_b200 import _lib, ops
L = _lib.lib(); dev = torch.device("cuda:0")
N, Dh, V = 8192, 512, 8192
h = torch.randn(N, Dh, device=dev); W = torch.randn(V, Dh, device=dev) * 0.04; b = torch.zeros(V, device=dev)
labels = torch.randint(0, V, (N,), device=dev)
rows = torch.from_numpy(np.flatnonzero(np.random.default_rng(0).random(N) < 0.15).astype(np.int32)).to(dev)
head = ops.PreparedHead(V, Dh, dev).prepare(W, b)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
names = ["lse", "dlogits", "dW", "dh"]
for rep in range(2):
    tl = torch.zeros(4 * 8192, dtype=torch.int64, device=dev)
    flush.zero_(); torch.cuda.synchronize()
    L.pero_debug_set_timeline(tl.data_ptr(), tl.numel() // 8192)
    loss_sum, lse, ws = ops.masked_ce_fwd(h, rows, labels, head)
    ops.masked_ce_bwd(h, rows, labels, head, lse, None, 1.0 / rows.numel(), ws=ws)
    torch.cuda.synchronize()
    L.pero_debug_set_timeline(None, 0)
t = tl.view(4, 8192).cpu()
for gi, nm in enumerate(names):
    c = t[gi][4096:4096 + 148 * 4].view(148, 4).double()
    used = c[:, 0] > 0
    c = c[used]; t0 = c[:, 0].min()
    ent, setup, epi, ex = [(c[:, i] - t0) / 1e3 for i in range(4)]
    u = t[gi][:4096].view(512, 8)
    nunits = int((u[:, 0] > 0).sum())
    print(f"{nm:8s}: ctas {int(used.sum())} | setup done {setup.min():.1f}..{setup.max():.1f} us | epi done {epi[epi>0].min():.1f}..{epi.max():.1f} | exit {ex.min():.1f}..{ex.max():.1f} | worker0 units {nunits}")
    base = int(u[0][6])
    for k in range(min(nunits, 4)):
        print("      unit", k, [int(v) - base if int(v) else -1 for v in u[k]])
