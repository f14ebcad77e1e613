"""Times the distance GEMM + arg-min kernel alone (pero_vq_assign_bf16) at the BASELINE shapes.
Usage: python scratch/gemm_perf.py [variant]   (PERO_ASSIGN_VARIANT: bit0 pairs, bit1 resident A)"""
import os
import sys

import torch

if len(sys.argv) > 1:
    os.environ["PERO_ASSIGN_VARIANT"] = sys.argv[1]
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from pero_pretraining_b200 import _lib, ops  # noqa: E402

L = _lib.lib()
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream().cuda_stream
print("variant", os.environ.get("PERO_ASSIGN_VARIANT", "default"))
for name, N, K, D in [("c1", 1024, 4096, 512), ("c2", 8192, 8192, 256), ("c4", 65536, 16384, 512), ("c5/8", 1 << 20, 8192, 512),
                      ("sq", 16384, 16384, 256)]:
    g = torch.Generator(device="cpu").manual_seed(1)
    w = torch.randn(K, D, generator=g).to(dev)
    xb = torch.randn(N, D, device=dev).bfloat16()
    cb = ops.PreparedCodebook(K, D, dev).prepare(w)
    packed = torch.empty(N, dtype=torch.int64, device=dev)
    ts = []
    for i in range(13):
        L.pero_vq_packed_init(packed.data_ptr(), N, stream)
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = L.pero_vq_assign_bf16(xb.data_ptr(), N, K, D, cb.blob.data_ptr(), 0, packed.data_ptr(), stream)
        e1.record()
        _lib.check(rc, "assign")
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    ms = sum(ts) / len(ts)
    # correctness spot check against fp32 torch on a slice
    idx, _ = ops.vq_unpack(packed)
    sl = slice(0, min(N, 2048))
    ref = torch.argmin((w * w).sum(1)[None, :] - 2 * xb[sl].float() @ w.t(), dim=1)
    agree = (idx[sl] == ref).float().mean().item()
    print(f"{name}: N={N} K={K} D={D}: {ms*1e3:8.1f} us (min {min(ts)*1e3:.1f})  {2.0*N*K*D/ms/1e9:7.1f} TFLOP/s  agree {agree:.4f}")
