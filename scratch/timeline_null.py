"""Per-unit stamps of the GEMM core with the epilogue that never reads TMEM (pero_gemm_tn_bf16 variant 16: dev build)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from pero_pretraining_b200 import _lib
L = _lib.lib(); dev = torch.device("cuda:0")
stream = torch.cuda.current_stream().cuda_stream
N, K, D = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (8192, 8192, 256))]
a = torch.randn(N, D, device=dev).bfloat16(); b = torch.randn(K, D, device=dev).bfloat16()
tl = torch.zeros(8192 + 148 * 4 + 64, dtype=torch.int64, device=dev)
for _ in range(3):
    tl.zero_()
    _lib.check(L.pero_gemm_tn_bf16(a.data_ptr(), N, b.data_ptr(), K, D, 16 + 3, 1, tl.data_ptr(), stream), "dbg")
    torch.cuda.synchronize()
t = tl[:4096].view(512, 8).cpu(); t0 = int(t[0][0])
for u in range(15):
    if int(t[u].max()) == 0: break
    print(f"{u:4d} " + " ".join(f"{(int(v)-t0) if int(v) else -1:11d}" for v in t[u]))
