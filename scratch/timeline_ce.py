"""Per-unit stamps (worker 0) and per-CTA wall-clock stamps of the masked-CE GEMMs at the bench shape (dev build)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from pero_pretraining_b200 import _lib, ops
L = _lib.lib(); dev = torch.device("cuda:0")
N, Dh, V = 8192, 512, 8192
h = torch.randn(N, Dh, device=dev); W = torch.randn(V, Dh, device=dev) * 0.04; b = torch.zeros(V, device=dev)
labels = torch.randint(0, V, (N,), device=dev)
rows = torch.from_numpy(np.flatnonzero(np.random.default_rng(0).random(N) < 0.15).astype(np.int32)).to(dev)
head = ops.PreparedHead(V, Dh, dev).prepare(W, b)
SLOTS = 4
tl = torch.zeros(8192 * SLOTS, dtype=torch.int64, device=dev)
for it in range(3):
    tl.zero_()
    torch.cuda.synchronize()
    L.pero_debug_set_timeline(tl.data_ptr(), SLOTS)
    loss_sum, lse, ws = ops.masked_ce_fwd(h, rows, labels, head, keep_logits=True)
    ops.masked_ce_bwd(h, rows, labels, head, lse, None, 1.0 / rows.numel(), ws=ws, ws_from_fwd=True, logits_in_ws=True)
    torch.cuda.synchronize()
    L.pero_debug_set_timeline(None, 0)
names = ["mma_top", "tempty_ok", "first_full", "issued", "epi_tfull", "epi_done", "prod_first", "prod_last"]
for s in range(SLOTS):
    t = tl[s * 8192: s * 8192 + 4096].view(512, 8).cpu()
    if int(t.max()) == 0: continue
    nz = t[0][t[0] > 0]
    t0 = int(nz.min())
    print(f"== GEMM launch {s}: worker 0 per-unit stamps (cycles)")
    print("unit " + " ".join(f"{n:>10s}" for n in names))
    for u in range(8):
        if int(t[u].max()) == 0: break
        print(f"{u:4d} " + " ".join(f"{(int(v)-t0) if int(v) else -1:10d}" for v in t[u]))
    c = tl[s * 8192 + 4096: s * 8192 + 4096 + 148 * 4].view(148, 4).cpu().numpy()
    live = c[:, 0] > 0
    if live.any():
        base = c[live, 0].min()
        rel = (c[live] - base) / 1000.0
        print(f"   CTAs {int(live.sum())}: start us min/max {rel[:,0].min():.1f}/{rel[:,0].max():.1f}  setup-done max {rel[:,1].max():.1f}  "
              f"epi-done min/mean/max {rel[:,2].min():.1f}/{rel[:,2].mean():.1f}/{rel[:,2].max():.1f}  exit max {rel[:,3].max():.1f}")
