"""Per-CTA wall-clock profile of the assign kernel at c2: entry / setup done / epilogue done / exit (ns)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from pero_pretraining_b200 import _lib, ops
L = _lib.lib(); dev = torch.device("cuda:0")
stream = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for N, K, D in [(8192, 8192, 256), (16384, 16384, 256)]:
    w = torch.randn(K, D, device=dev); xb = torch.randn(N, D, device=dev).bfloat16()
    cb = ops.PreparedCodebook(K, D, dev).prepare(w)
    packed = torch.empty(N, dtype=torch.int64, device=dev)
    tl = torch.zeros(4096 + 148 * 4 + 64, dtype=torch.int64, device=dev)
    L.pero_debug_set_timeline(tl.data_ptr(), tl.numel() // 8192)
    for cold in (0, 1, 1):
        tl.zero_(); L.pero_vq_packed_init(packed.data_ptr(), N, stream)
        if cold: flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.pero_vq_assign_bf16(xb.data_ptr(), N, K, D, cb.blob.data_ptr(), 0, packed.data_ptr(), stream), "assign")
        e1.record(); torch.cuda.synchronize()
        t = tl[4096:4096 + 148 * 4].view(148, 4).cpu().double()
        t0 = t[:, 0].min()
        ent, setup, epi, ex = [(t[:, i] - t0) / 1e3 for i in range(4)]
        print(f"N={N} K={K} cold={cold}: event {e0.elapsed_time(e1)*1e3:.1f} us | entry spread {ent.max():.1f} us | setup done {setup.min():.1f}..{setup.max():.1f} | "
              f"epilogue done {epi[epi>0].min():.1f}..{epi.max():.1f} | exit {ex.min():.1f}..{ex.max():.1f} | per-cta busy {(ex-ent).min():.1f}..{(ex-ent).max():.1f}")
    L.pero_debug_set_timeline(None, 0)
