"""Out-of-bounds WRITE check of the C-ABI entry points (compute-sanitizer is closed on this GPU pool, so the library is
checked with guard bands of its own): every output buffer, blob and workspace handed to the library sits between two
4 KiB bands of a known pattern inside one larger allocation, is sized EXACTLY as the *_bytes() query / the documented
shape says, and the bands must be untouched afterwards.  Ragged shapes on purpose (nothing a multiple of a tile)."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GUARD = 4096
PATTERN = 0xA5


class Guarded:
    def __init__(self, dev):
        self.dev, self.items = dev, []

    def buf(self, nbytes, dtype=torch.uint8, fill=None):
        nbytes = int(nbytes)
        total = GUARD + ((nbytes + 255) // 256) * 256 + 256 + GUARD
        raw = torch.full((total,), PATTERN, dtype=torch.uint8, device=self.dev)
        off = GUARD + (-(raw.data_ptr() + GUARD) % 256)          # 256-byte aligned payload
        view = raw[off:off + nbytes]
        if fill is not None:
            view.fill_(fill)
        self.items.append((raw, off, nbytes))
        esz = torch.empty((), dtype=dtype).element_size()
        return view.view(dtype) if nbytes % esz == 0 and nbytes else view

    def check(self, what):
        torch.cuda.synchronize()
        for i, (raw, off, n) in enumerate(self.items):
            assert bool((raw[:off] == PATTERN).all()), f"{what}: buffer {i} ({n} bytes) written BEFORE its start"
            assert bool((raw[off + n:] == PATTERN).all()), f"{what}: buffer {i} ({n} bytes) written PAST its end"


def _s():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("n_lines,frames,K,D,cf", [(3, 37, 300, 72, 1), (1, 1, 5, 8, 1), (5, 129, 1000, 200, 0), (2, 700, 17000, 40, 1)])
def test_quantizer_forward_stays_inside_its_buffers(cuda_dev, n_lines, frames, K, D, cf):
    from pero_pretraining_b200 import _lib
    L = _lib.lib()
    N = n_lines * frames
    g = Guarded(cuda_dev)
    gen = torch.Generator(device="cpu").manual_seed(N + K)
    x = torch.randn(n_lines, D, frames, generator=gen).to(cuda_dev) if cf else torch.randn(N, D, generator=gen).to(cuda_dev)
    w = g.buf(K * D * 4, torch.float32); w.copy_(torch.randn(K * D, generator=gen))
    ema_w = g.buf(K * D * 4, torch.float32); ema_w.copy_(w)
    cs = g.buf(K * 4, torch.float32, fill=0); cs.fill_(1.0)
    cbb = L.pero_vq_codebook_bytes(K, D)
    cb = g.buf(cbb)
    _lib.check(L.pero_vq_codebook_prepare(w.data_ptr(), K, D, cb.data_ptr(), cbb, _s()), "prepare")
    out = g.buf(N * D * 4, torch.float32)
    idx = g.buf(N * 8, torch.int64)
    for upd in (1, 0):
        wsb = L.pero_vq_forward_workspace_bytes(N, K, D, upd)
        ws = g.buf(wsb)
        _lib.check(L.pero_vq_forward(x.data_ptr(), n_lines, frames, cf, K, D, cb.data_ptr(), cbb, w.data_ptr(), ema_w.data_ptr(),
                                     cs.data_ptr(), 0.99, 1e-5, upd, out.data_ptr(), idx.data_ptr(), ws.data_ptr(), wsb, _s()), "forward")
        g.check(f"pero_vq_forward update={upd}")
        assert int(idx.min()) >= 0 and int(idx.max()) < K
    # collapsed assignment: the long-segment kernels of the EMA sums
    idx.fill_(K - 1)
    sums = g.buf((K * D + K) * 4, torch.float32)
    wsb = L.pero_vq_ema_workspace_bytes(N, K, D)
    ws = g.buf(wsb)
    xr = torch.randn(N, D, generator=gen).to(cuda_dev)
    _lib.check(L.pero_vq_ema_accumulate(xr.data_ptr(), idx.data_ptr(), N, K, D, sums.data_ptr(), ws.data_ptr(), wsb, _s()), "acc")
    g.check("pero_vq_ema_accumulate collapsed")
    counts = g.buf(K * 8, torch.int64)
    _lib.check(L.pero_vq_counts(idx.data_ptr(), N, K, counts.data_ptr(), _s()), "counts")
    packed = g.buf(N * 8, torch.int64)
    _lib.check(L.pero_vq_packed_init(packed.data_ptr(), N, _s()), "init")
    dmin = g.buf(N * 4, torch.float32)
    _lib.check(L.pero_vq_unpack(packed.data_ptr(), N, idx.data_ptr(), dmin.data_ptr(), _s()), "unpack")
    g.check("counts / packed")


@pytest.mark.parametrize("N,Dh,V,p,bf16", [(150, 96, 700, 0.3, 0), (333, 128, 257, 0.5, 1), (90, 576, 300, 0.4, 0), (64, 64, 64, 1.0, 0)])
def test_masked_ce_stays_inside_its_buffers(cuda_dev, N, Dh, V, p, bf16):
    from pero_pretraining_b200 import _lib
    L = _lib.lib()
    g = Guarded(cuda_dev)
    gen = torch.Generator(device="cpu").manual_seed(N + V)
    h = torch.randn(N, Dh, generator=gen).to(cuda_dev)
    if bf16:
        h = h.bfloat16()
    W = torch.randn(V, Dh, generator=gen).to(cuda_dev) * 0.05
    b = torch.randn(V, generator=gen).to(cuda_dev) * 0.05
    labels = torch.randint(0, V, (N,), generator=gen).to(cuda_dev)
    rows_np = np.flatnonzero(np.random.default_rng(N).random(N) < p).astype(np.int32)
    if rows_np.size == 0:
        rows_np = np.array([0], dtype=np.int32)
    rows = torch.from_numpy(rows_np).to(cuda_dev)
    M = int(rows.numel())
    hb = L.pero_head_bytes(V, Dh)
    head = g.buf(hb)
    _lib.check(L.pero_head_prepare(W.data_ptr(), b.data_ptr(), V, Dh, head.data_ptr(), hb, _s()), "head")
    wsb = L.pero_masked_ce_workspace_bytes(N, M, V, Dh)
    ws = g.buf(wsb)
    loss = g.buf(4, torch.float32)
    lse = g.buf(M * 4, torch.float32)
    _lib.check(L.pero_masked_ce_fwd(h.data_ptr(), bf16, N, Dh, rows.data_ptr(), M, labels.data_ptr(), head.data_ptr(), V,
                                    loss.data_ptr(), lse.data_ptr(), ws.data_ptr(), wsb, _s()), "fwd")
    g.check("pero_masked_ce_fwd")
    d_h = g.buf(N * Dh * (2 if bf16 else 4), torch.bfloat16 if bf16 else torch.float32)
    d_W = g.buf(V * Dh * 4, torch.float32)
    d_b = g.buf(V * 4, torch.float32)
    _lib.check(L.pero_masked_ce_bwd(None, bf16, N, Dh, rows.data_ptr(), M, labels.data_ptr(), head.data_ptr(), V, None, None,
                                    1.0 / M, d_h.data_ptr(), d_W.data_ptr(), d_b.data_ptr(), ws.data_ptr(), wsb, _s()), "bwd")
    g.check("pero_masked_ce_bwd on the forward's workspace")
    # the two-phase backward of the data-parallel schedule and a fresh-workspace backward give the same gradients
    ref = (d_h.clone(), d_W.clone(), d_b.clone())
    d_h.zero_(); d_W.zero_(); d_b.zero_()
    _lib.check(L.pero_masked_ce_bwd_range(None, bf16, N, Dh, rows.data_ptr(), M, labels.data_ptr(), head.data_ptr(), V, None, None,
                                          1.0 / M, 0, V, None, d_W.data_ptr(), None, ws.data_ptr(), wsb, _s()), "phase 1")
    _lib.check(L.pero_masked_ce_bwd_range(None, bf16, N, Dh, rows.data_ptr(), M, labels.data_ptr(), head.data_ptr(), V, None, None,
                                          1.0 / M, 0, V, d_h.data_ptr(), None, d_b.data_ptr(), ws.data_ptr(), wsb, _s()), "phase 2")
    g.check("two-phase backward")
    for got, want in zip((d_h, d_W, d_b), ref):
        assert torch.equal(got, want)
    ws2 = g.buf(wsb)
    _lib.check(L.pero_masked_ce_bwd(h.data_ptr(), bf16, N, Dh, rows.data_ptr(), M, labels.data_ptr(), head.data_ptr(), V,
                                    lse.data_ptr(), None, 1.0 / M, d_h.data_ptr(), d_W.data_ptr(), d_b.data_ptr(), ws2.data_ptr(),
                                    wsb, _s()), "bwd fresh")
    g.check("pero_masked_ce_bwd with its own gather")
    for got, want in zip((d_h, d_W, d_b), ref):
        assert torch.allclose(got.float(), want.float(), rtol=2e-3, atol=2e-5 * float(want.float().abs().max()) + 1e-7)
    errors = g.buf(3 * 8, torch.int64)
    rank = g.buf(M * 4, torch.int32)
    ks = (ctypes.c_int32 * 3)(1, 3, 10)
    _lib.check(L.pero_masked_ce_eval(h.data_ptr(), bf16, N, Dh, rows.data_ptr(), M, labels.data_ptr(), head.data_ptr(), V,
                                     ctypes.addressof(ks), 3, loss.data_ptr(), lse.data_ptr(), rank.data_ptr(), errors.data_ptr(),
                                     ws.data_ptr(), wsb, _s()), "eval")
    g.check("pero_masked_ce_eval")


def test_small_kernels_stay_inside_their_buffers(cuda_dev):
    from pero_pretraining_b200 import _lib
    L = _lib.lib()
    g = Guarded(cuda_dev)
    gen = torch.Generator(device="cpu").manual_seed(3)
    n = 100003
    a = torch.randn(n, generator=gen).to(cuda_dev); b = torch.randn(n, generator=gen).to(cuda_dev)
    out = g.buf(4, torch.float32)
    wsb = L.pero_mse_workspace_bytes(n)
    ws = g.buf(wsb)
    _lib.check(L.pero_mse_fwd(a.data_ptr(), b.data_ptr(), n, 1.0, 0.25, out.data_ptr(), ws.data_ptr(), wsb, _s()), "mse")
    ga, gb = g.buf(n * 4, torch.float32), g.buf(n * 4, torch.float32)
    _lib.check(L.pero_mse_bwd(a.data_ptr(), b.data_ptr(), n, 0.5, None, ga.data_ptr(), gb.data_ptr(), _s()), "mse bwd")
    gx = g.buf(n * 4, torch.float32)
    _lib.check(L.pero_vq_st_commit_bwd(a.data_ptr(), b.data_ptr(), a.data_ptr(), n, 0.5, None, gx.data_ptr(), _s()), "st")
    g.check("mse / straight-through kernels")
    # pixel masking: ragged width, last partial column
    Nl, C, H, W, pw = 2, 3, 40, 77, 8
    x = g.buf(Nl * C * H * W * 4, torch.float32); x.copy_(torch.rand(Nl * C * H * W, generator=gen))
    tile = torch.rand(C, H, pw, generator=gen).to(cuda_dev)
    rows = torch.tensor([0, 9, 10, 19], dtype=torch.int32, device=cuda_dev)          # frames 9 and 19 are the partial ones
    _lib.check(L.pero_mask_pixels(x.data_ptr(), Nl, C, H, W, rows.data_ptr(), 4, 10, pw, tile.data_ptr(), _s()), "pixels")
    g.check("pero_mask_pixels")
    V, Dh = 300, 72
    cbb = L.pero_vq_codebook_bytes(V, Dh)
    cb = g.buf(cbb)
    Wt = torch.randn(V, Dh, generator=gen).to(cuda_dev)
    _lib.check(L.pero_head_argmax_prepare(Wt.data_ptr(), None, V, Dh, cb.data_ptr(), cbb, _s()), "argmax prepare")
    N = 777
    mask = (torch.rand(N, generator=gen) < 0.3).to(torch.uint8).to(cuda_dev)
    rows_out = g.buf(N * 4, torch.int32)
    count = g.buf(4, torch.int32)
    wsb = L.pero_mask_compact_workspace_bytes(N)
    ws = g.buf(wsb)
    _lib.check(L.pero_mask_compact(mask.data_ptr(), 2, 1, None, N, rows_out.data_ptr(), count.data_ptr(), ws.data_ptr(), wsb, _s()), "compact")
    g.check("argmax prepare / mask compaction")
    assert int(count) == int(mask.sum())
