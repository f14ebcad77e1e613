"""MMA/TMA ceiling of the GEMM core: epilogue that never reads TMEM (bit2) or only reads it (bit3)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from pero_pretraining_b200 import _lib
L = _lib.lib(); dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream().cuda_stream
for name, N, K, D in [("c2", 8192, 8192, 256), ("sq", 16384, 16384, 256), ("c4", 65536, 16384, 512)]:
    a = torch.randn(N, D, device=dev).bfloat16(); b = torch.randn(K, D, device=dev).bfloat16()
    out = torch.zeros(N, device=dev)
    for variant, tag in [(4 + 3, "null pair+res"), (8 + 3, "load pair+res"), (4 + 2, "null 1cta+res"), (8 + 2, "load 1cta+res"), (4 + 1, "null pair stream"), (4 + 0, "null 1cta stream")]:
        ts = []
        for i in range(9):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = L.pero_debug_gemm_tn(a.data_ptr(), N, b.data_ptr(), K, D, variant, 1, out.data_ptr(), stream)
            e1.record(); _lib.check(rc, "dbg"); torch.cuda.synchronize()
            if i >= 3: ts.append(e0.elapsed_time(e1))
        ms = sum(ts) / len(ts)
        print(f"{name} {tag:18s}: {ms*1e3:8.1f} us  {2.0*N*K*D/ms/1e9:7.1f} TFLOP/s")
