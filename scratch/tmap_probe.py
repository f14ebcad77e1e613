import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from pero_pretraining_b200 import _lib, ops
L = _lib.lib(); dev = torch.device("cuda:0")
for shape in [(96, 32, 64), (27, 32, 128), (128, 32, 64), (96, 128, 64), (256, 128, 64), (96, 64, 64), (300, 100, 64)]:
    for variant in (0, 1):
        for splits in (1, 2):
            ra, rb, kd = shape
            a = torch.randn(ra, kd, device=dev).bfloat16(); b = torch.randn(rb, kd, device=dev).bfloat16()
            out = torch.zeros(splits, ra, rb, device=dev)
            rc = L.pero_debug_gemm_tn(a.data_ptr(), ra, b.data_ptr(), rb, kd, variant, splits, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            err = float((out.sum(0).double() - a.double() @ b.double().t()).abs().max()) if rc == 0 else None
            print(shape, "variant", variant, "splits", splits, "rc", rc, "err", err)
