"""Two launches of the distance GEMM + arg-min kernel per shape (c2, c4) for an ncu capture."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from pero_pretraining_b200 import _lib, ops
L = _lib.lib(); dev = torch.device("cuda:0")
stream = torch.cuda.current_stream().cuda_stream
for name, N, K, D in [("c2", 8192, 8192, 256), ("sq", 16384, 16384, 256), ("c4", 65536, 16384, 512)]:
    g = torch.Generator(device="cpu").manual_seed(1)
    w = torch.randn(K, D, generator=g).to(dev)
    xb = torch.randn(N, D, device=dev).bfloat16()
    cb = ops.PreparedCodebook(K, D, dev).prepare(w)
    packed = torch.empty(N, dtype=torch.int64, device=dev)
    for i in range(2):
        L.pero_vq_packed_init(packed.data_ptr(), N, stream)
        _lib.check(L.pero_vq_assign_bf16(xb.data_ptr(), N, K, D, cb.blob.data_ptr(), 0, packed.data_ptr(), stream), "assign")
    torch.cuda.synchronize()
print("done")
