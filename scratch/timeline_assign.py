"""Timeline of worker 0 for the real assign kernel (ArgminEpi)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from pero_pretraining_b200 import _lib, ops
L = _lib.lib(); dev = torch.device("cuda:0")
stream = torch.cuda.current_stream().cuda_stream
names = ["mma_top", "mma_tempty_ok", "mma_first_full", "mma_issued", "epi_tfull_ok", "epi_done", "prod_first", "prod_last"]
for N, K, D in [(16384, 16384, 256)]:
    w = torch.randn(K, D, device=dev); xb = torch.randn(N, D, device=dev).bfloat16()
    cb = ops.PreparedCodebook(K, D, dev).prepare(w)
    packed = torch.empty(N, dtype=torch.int64, device=dev)
    tl = torch.zeros(256 * 8, dtype=torch.int64, device=dev)
    L.pero_debug_set_timeline(tl.data_ptr(), tl.numel() // 8192)
    for _ in range(2):
        tl.zero_(); L.pero_vq_packed_init(packed.data_ptr(), N, stream)
        _lib.check(L.pero_vq_assign_bf16(xb.data_ptr(), N, K, D, cb.blob.data_ptr(), 0, packed.data_ptr(), stream), "assign")
        torch.cuda.synchronize()
    L.pero_debug_set_timeline(None, 0)
    t = tl.view(256, 8).cpu(); t0 = int(t[0][6])
    print("unit " + " ".join(f"{n:>14s}" for n in names))
    for u in list(range(6)) + list(range(40, 46)):
        print(f"{u:4d} " + " ".join(f"{(int(v)-t0) if int(v) else -1:14d}" for v in t[u]))
