"""First GPU probe of the tcgen05 GEMM core: full-matrix check vs torch, then assign vs fp64 brute force.
Usage: python scratch/gpu_probe1.py <variant 0..3>
"""
import ctypes
import os
import sys
import time

import torch

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
os.environ["PERO_ASSIGN_VARIANT"] = str(variant)
L = ctypes.CDLL(os.path.join(os.path.dirname(__file__), "..", "pero_pretraining_b200", "libpero_b200.so"))
vp, i64, ci, sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_size_t
L.pero_strerror.restype = ctypes.c_char_p
L.pero_debug_gemm_tn.argtypes = [vp, i64, vp, i64, i64, ci, ci, vp, vp]
L.pero_vq_codebook_bytes.restype = sz
L.pero_vq_codebook_bytes.argtypes = [i64, i64]
L.pero_vq_codebook_prepare.argtypes = [vp, i64, i64, vp, sz, vp]
L.pero_vq_assign_workspace_bytes.restype = sz
L.pero_vq_assign_workspace_bytes.argtypes = [i64, i64, i64]
L.pero_vq_assign.argtypes = [vp, i64, i64, ci, i64, i64, vp, i64, vp, vp, vp, vp, vp, sz, vp]


def chk(rc, what):
    if rc != 0:
        print(f"FAIL {what}: rc={rc} {L.pero_strerror(rc).decode()}")
        sys.exit(2)


dev = torch.device("cuda:0")
print("device", torch.cuda.get_device_name(0), "variant", variant, "check_device", L.pero_check_device())
stream = torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)


def gemm_case(ra, rb, kd, splits=1):
    a = torch.randn(ra, kd, device=dev).bfloat16()
    b = torch.randn(rb, kd, device=dev).bfloat16()
    out = torch.full((max(splits, 1), ra, rb), float("nan"), device=dev)
    rc = L.pero_debug_gemm_tn(a.data_ptr(), ra, b.data_ptr(), rb, kd, variant, splits, out.data_ptr(), stream)
    chk(rc, f"gemm {ra}x{rb}x{kd}")
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t()
    got = out.sum(0)
    err = (got - ref).abs().max().item()
    bad = (~torch.isfinite(got)).sum().item()
    print(f"gemm {ra}x{rb}x{kd} splits={splits}: max|err|={err:.3e} nonfinite={bad} ref_absmax={ref.abs().max().item():.1f}")
    if bad or err > 1e-2 * max(1.0, ref.abs().max().item()):
        rows = ((got - ref).abs() > 1e-2).any(1).nonzero().flatten()[:16].tolist()
        cols = ((got - ref).abs() > 1e-2).any(0).nonzero().flatten()[:16].tolist()
        print("  first bad rows", rows, "cols", cols)
        print("  got[0,:8]", got[0, :8].tolist(), "\n  ref[0,:8]", ref[0, :8].tolist())
        return False
    return True


ok = True
resident = bool(variant & 2)
ok &= gemm_case(128, 256, 64)
ok &= gemm_case(128, 256, 256)
ok &= gemm_case(300, 700, 256)
ok &= gemm_case(1000, 1000, 512)
if not resident:
    ok &= gemm_case(512, 512, 2048)
    ok &= gemm_case(300, 520, 1280, splits=3)
if not ok:
    print("GEMM CHECK FAILED")
    sys.exit(1)

# timing of the big plain GEMM
ra = rb = 8192
for kd in (256, 512):
    a = torch.randn(ra, kd, device=dev).bfloat16()
    b = torch.randn(rb, kd, device=dev).bfloat16()
    out = torch.empty(ra, rb, device=dev)
    for _ in range(3):
        chk(L.pero_debug_gemm_tn(a.data_ptr(), ra, b.data_ptr(), rb, kd, variant, 1, out.data_ptr(), stream), "big")
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10):
        L.pero_debug_gemm_tn(a.data_ptr(), ra, b.data_ptr(), rb, kd, variant, 1, out.data_ptr(), stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"store-gemm 8192x8192x{kd}: {ms*1e3:.1f} us  {2*ra*rb*kd/ms/1e9:.1f} TFLOP/s")


def assign_case(nl, T, K, D, iters=0):
    x = torch.randn(nl, D, 1, T, device=dev)
    w = torch.randn(K, D, device=dev)
    N = nl * T
    cbb = L.pero_vq_codebook_bytes(K, D)
    cb = torch.empty(cbb, dtype=torch.uint8, device=dev)
    chk(L.pero_vq_codebook_prepare(w.data_ptr(), K, D, cb.data_ptr(), cbb, stream), "cb prepare")
    wsb = L.pero_vq_assign_workspace_bytes(N, K, D)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    idx = torch.full((N,), -1, dtype=torch.int64, device=dev)
    dmin = torch.zeros(N, device=dev)
    xr = torch.empty(N, D, device=dev)

    def run():
        return L.pero_vq_assign(x.data_ptr(), nl, T, 1, K, D, cb.data_ptr(), 0, idx.data_ptr(), dmin.data_ptr(), None,
                                xr.data_ptr(), ws.data_ptr(), wsb, stream)

    chk(run(), f"assign {N}x{K}x{D}")
    torch.cuda.synchronize()
    flat = x.permute(0, 2, 3, 1).reshape(N, D)
    print(f"  x_rows copy exact: {torch.equal(xr, flat)}")
    # fp64 brute force in row chunks
    ref_idx = torch.empty(N, dtype=torch.int64, device=dev)
    gap = torch.empty(N, dtype=torch.float64, device=dev)
    w64 = w.double()
    wn = (w64 * w64).sum(1)
    for s in range(0, N, 4096):
        f = flat[s:s + 4096].double()
        d = (f * f).sum(1, keepdim=True) + wn - 2 * f @ w64.t()
        top2 = torch.topk(d, 2, dim=1, largest=False)
        ref_idx[s:s + 4096] = top2.indices[:, 0]
        gap[s:s + 4096] = (top2.values[:, 1] - top2.values[:, 0]) / top2.values[:, 0].abs().clamp_min(1e-30)
    mism = idx != ref_idx
    nm = int(mism.sum())
    print(f"assign N={N} K={K} D={D}: mismatches {nm}/{N} ({100.0*nm/N:.3f}%), "
          f"max rel gap among mismatches {gap[mism].max().item() if nm else 0:.2e}, median gap {gap.median().item():.2e}, "
          f"idx range [{idx.min().item()},{idx.max().item()}]")
    if iters:
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(iters):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"  assign (prep+gemm+unpack) {ms*1e3:.1f} us/iter  gemm-equivalent {2*N*K*D/ms/1e9:.1f} TFLOP/s")
    return nm / N


r = assign_case(8, 128, 4096, 512, iters=20)
r = max(r, assign_case(64, 128, 8192, 256, iters=20))
r = max(r, assign_case(7, 100, 1000, 200))
r = max(r, assign_case(512, 128, 16384, 512, iters=10))
print("PROBE_OK" if r < 0.02 else "PROBE_MISMATCH_TOO_HIGH")
