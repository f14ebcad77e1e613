"""Host-side time of each segment of the module-API step (the bench's e2e leg), to see what bounds it."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench
from pero_pretraining_b200 import LinearHead, VectorQuantizer
c = bench.CFG; dev = torch.device("cuda:0")
batch = bench.make_batch(0)
vq = VectorQuantizer(c["K"], c["D"], c["commitment_cost"], c["decay"], c["epsilon"]).to(dev).train()
head = LinearHead(c["Dh"], c["V"]).to(dev)
with torch.no_grad():
    vq.embedding.weight.copy_(batch["weight"]); vq.ema_w.copy_(batch["weight"]); vq.ema_cluster_size.fill_(1.0)
    head.linear.weight.copy_(batch["W"]); head.linear.bias.copy_(batch["b"])
x_host, h_host = batch["x"].pin_memory(), batch["h"].pin_memory()
gq = batch["gq"].to(dev); mask = batch["mask"]
xd, hd = x_host.to(dev), h_host.to(dev)
def seg_times(sync_each):
    T = {}
    def mark(name, t0):
        if sync_each: torch.cuda.synchronize()
        T[name] = T.get(name, 0.0) + time.perf_counter() - t0
    for it in range(25):
        if it == 5: T.clear()
        t0 = time.perf_counter(); xd.copy_(x_host, non_blocking=True); hd.copy_(h_host, non_blocking=True); mark("h2d", t0)
        x = xd.detach().requires_grad_(True); h = hd.detach().requires_grad_(True)
        t0 = time.perf_counter(); q, idx = vq(x); mark("vq.forward", t0)
        t0 = time.perf_counter(); lc = vq.calculate_loss(q, x); mark("calculate_loss", t0)
        t0 = time.perf_counter(); lm = head.masked_loss(h, idx.view(c["lines"], c["frames"]), mask, None, None); mark("masked_loss", t0)
        t0 = time.perf_counter(); loss = lc + lm; head.linear.weight.grad = None; head.linear.bias.grad = None; mark("add", t0)
        t0 = time.perf_counter(); torch.autograd.backward([loss, q], [None, gq]); mark("backward", t0)
        t0 = time.perf_counter(); v = float(loss.item()); mark("item", t0)
    return {k: round(v / 20 * 1e6, 1) for k, v in T.items()}
print("async  (host launch time per segment, us):", seg_times(False))
print("synced (host + device time per segment, us):", seg_times(True))
