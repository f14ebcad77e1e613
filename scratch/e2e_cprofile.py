import os, sys, time, cProfile, pstats, numpy as np, torch
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
gq = batch["gq"].to(dev); mask = batch["mask"]
xd, hd = batch["x"].to(dev), batch["h"].to(dev).bfloat16()
vq.enable_cuda_graph()
torch.autograd.set_multithreading_enabled(False)
def step():
    x = xd.detach().requires_grad_(True); h = hd.detach().requires_grad_(True)
    q, idx = vq(x)
    loss = vq.calculate_loss(q, x) + head.masked_loss(h, idx.view(c["lines"], c["frames"]), mask, None, None)
    head.linear.weight.grad = None; head.linear.bias.grad = None
    torch.autograd.backward([loss, q], [None, gq])
    return float(loss.item())
for _ in range(10): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): step()
torch.cuda.synchronize()
print("us/step without profiler:", (time.perf_counter() - t0) / 50 * 1e6)
pr = cProfile.Profile(); pr.enable()
for _ in range(50): step()
pr.disable()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(70)
