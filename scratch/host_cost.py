"""Host-side cost (us per call, no device sync inside the loop) of each tensor-level op at the bench shape."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench
from pero_pretraining_b200 import ops, _lib
c = bench.CFG; dev = torch.device("cuda:0"); L = _lib.lib()
b = bench.make_batch(0)
x = b["x"].to(dev).view(c["lines"], c["D"], c["frames"]); h = b["h"].to(dev).view(-1, c["Dh"]); gq = b["gq"].to(dev).view_as(x)
W, bias, weight = b["W"].to(dev), b["b"].to(dev), b["weight"].to(dev)
ema_w, cs = weight.clone(), torch.ones(c["K"], device=dev)
rows = torch.from_numpy(np.flatnonzero(b["mask"].reshape(-1) == 1).astype(np.int32)).to(dev)
cb = ops.PreparedCodebook(c["K"], c["D"], dev).prepare(weight); head = ops.PreparedHead(c["V"], c["Dh"], dev).prepare(W, bias)
idx, _, x_rows = ops.vq_assign(x, cb, c["lines"], c["frames"], True, want_rows=True)
q = ops.vq_gather_st(x_rows, idx, weight, c["lines"], c["frames"], True)
sums = ops.vq_ema_accumulate(x_rows, idx, c["K"])
loss_sum, lse, ws = ops.masked_ce_fwd(h, rows, idx, head)
def t(name, fn, n=100):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n):
        fn()
        if i % 10 == 9: torch.cuda.synchronize()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n * 1e6
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n):
        fn(); torch.cuda.synchronize()
    ds = (time.perf_counter() - t0) / n * 1e6
    print(f"{name:28s} async {dt:7.1f} us   synced {ds:7.1f} us")
t("pero_version (ctypes)", lambda: L.pero_version())
t("torch.empty(1M f32)", lambda: torch.empty(1 << 20, device=dev))
t("current_stream", lambda: torch.cuda.current_stream().cuda_stream)
t("vq_assign", lambda: ops.vq_assign(x, cb, c["lines"], c["frames"], True, want_rows=True))
t("vq_gather_st", lambda: ops.vq_gather_st(x_rows, idx, weight, c["lines"], c["frames"], True))
t("vq_ema_accumulate", lambda: ops.vq_ema_accumulate(x_rows, idx, c["K"]))
t("vq_ema_apply", lambda: ops.vq_ema_apply(sums, ema_w, cs, weight, 0.99, 1e-5, cb))
t("mse_fwd", lambda: ops.mse_fwd(q, x, 0.0, 0.25))
t("mse_bwd", lambda: ops.mse_bwd(q, x, 1e-6, None, False, True))
t("st_commit_bwd", lambda: ops.vq_st_commit_bwd(gq, q, x, 1e-6))
t("head.prepare", lambda: head.prepare(W, bias))
t("masked_ce_fwd", lambda: ops.masked_ce_fwd(h, rows, idx, head))
t("masked_ce_bwd", lambda: ops.masked_ce_bwd(h, rows, idx, head, lse, None, 1e-3, ws=ws, ws_from_fwd=True))
a16 = torch.randn(1280, 512, device=dev).bfloat16(); b16 = torch.randn(512, 512, device=dev).bfloat16()
out = torch.zeros(1, 1280, 512, device=dev); packed = torch.empty(8192, dtype=torch.int64, device=dev)
s = torch.cuda.current_stream().cuda_stream
t("raw plain launch (packed_init)", lambda: L.pero_vq_packed_init(packed.data_ptr(), 8192, s), 300)
for v in (0, 1, 2, 3):
    t(f"raw GEMM launch variant {v}", lambda: L.pero_debug_gemm_tn(a16.data_ptr(), 1280, b16.data_ptr(), 512, 512, v, 1, out.data_ptr(), s), 300)
