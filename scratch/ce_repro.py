import os, sys, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from pero_pretraining_b200 import _lib, ops
L = _lib.lib(); dev = torch.device("cuda:0")
def run(N, Dh, V, M, seed=0):
    g = torch.Generator().manual_seed(seed)
    h = torch.randn(N, Dh, generator=g).to(dev); W = (torch.randn(V, Dh, generator=g) * 0.04).to(dev); b = torch.zeros(V, device=dev)
    labels = torch.randint(0, V, (N,), generator=g).to(dev)
    rows = torch.sort(torch.randperm(N, generator=g)[:M]).values.int().to(dev)
    head = ops.PreparedHead(V, Dh, dev).prepare(W, b)
    loss_sum, lse, ws = ops.masked_ce_fwd(h, rows, labels, head)
    for mode in ("reuse", "regather"):
        try:
            ops.masked_ce_bwd(h, rows, labels, head, lse, None, 1.0 / M, ws=ws, ws_from_fwd=(mode == "reuse"))
            torch.cuda.synchronize()
            print(N, Dh, V, M, mode, "ok")
        except Exception as e:
            print(N, Dh, V, M, mode, "FAIL", str(e)[:80])
for args in [(256, 512, 2048, 77), (1024, 512, 4096, 150), (96, 32, 96, 27), (185, 96, 1000, 90), (18, 64, 300, 18), (150, 512, 257, 3), (128, 64, 200, 38), (8192, 512, 8192, 1245)]:
    run(*args)
