"""Per-unit clock64 timeline of worker 0 of the GEMM core (null epilogue)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from pero_pretraining_b200 import _lib
L = _lib.lib(); dev = torch.device("cuda:0")
stream = torch.cuda.current_stream().cuda_stream
N, K, D = 8192, 8192, 256
a = torch.randn(N, D, device=dev).bfloat16(); b = torch.randn(K, D, device=dev).bfloat16()
names = ["mma_top", "mma_tempty_ok", "mma_first_full", "mma_issued", "epi_tfull_ok", "epi_done", "prod_first", "prod_last"]
for variant, tag in [(16 + 3, "pair+res"), (16 + 0, "1cta stream"), (16 + 2, "1cta res")]:
    tl = torch.zeros(64 * 8, dtype=torch.int64, device=dev)
    for _ in range(2):
        tl.zero_()
        _lib.check(L.pero_debug_gemm_tn(a.data_ptr(), N, b.data_ptr(), K, D, variant, 1, tl.data_ptr(), stream), "dbg")
        torch.cuda.synchronize()
    t = tl.view(64, 8).cpu()
    t0 = int(t[0][6]) if int(t[0][6]) else int(t[0][0])
    print(f"== {tag}: stamps relative to producer's first issue (cycles)")
    print("unit " + " ".join(f"{n:>14s}" for n in names))
    for u in range(16):
        if int(t[u].max()) == 0: break
        print(f"{u:4d} " + " ".join(f"{(int(v)-t0) if int(v) else -1:14d}" for v in t[u]))
