"""Masked-CE forward + backward at the bench shape on warm caches (for ncu --cache-control none)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from pero_pretraining_b200 import ops
dev = torch.device("cuda:0")
N, Dh, V = 8192, 512, 8192
h = torch.randn(N, Dh, device=dev); W = torch.randn(V, Dh, device=dev) * 0.04; b = torch.zeros(V, device=dev)
labels = torch.randint(0, V, (N,), device=dev)
rows = torch.from_numpy(np.flatnonzero(np.random.default_rng(0).random(N) < 0.15).astype(np.int32)).to(dev)
head = ops.PreparedHead(V, Dh, dev).prepare(W, b)
for _ in range(3):
    loss_sum, lse, ws = ops.masked_ce_fwd(h, rows, labels, head, keep_logits=True)
    ops.masked_ce_bwd(h, rows, labels, head, lse, None, 1.0 / rows.numel(), ws=ws, ws_from_fwd=True, logits_in_ws=True)
torch.cuda.synchronize()
print("done")
