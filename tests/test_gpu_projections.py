"""GPU parity of the 1x1 projections folded around the quantizer (models/autoencoders.py:114-115, 142-147):
the split-bf16 tensor-core projection against an fp64 matmul, the projected-codebook gather against torch indexing, the
fused VQVAE.quantize against the unfused module path (torch.nn.Conv2d around VectorQuantizer, TF32 off) including every
gradient, the golden VQVAE replay through the fused path, and guard bands around every buffer of the new entry points.

Tolerances: the projection splits every fp32 operand into bf16 hi + lo and drops the lo x lo product: per-element error
<= 4e-5 * sum_k |a_k b_k| (measured ~1e-5); bit-exact where nothing is computed (gather, bf16 copy == rounded fp32 rows,
packed reset)."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu
EPS_TIE = 2e-3
INT64_MAX = np.iinfo(np.int64).max


@pytest.mark.parametrize("n_lines,frames,C,D,cf,bias", [
    (64, 128, 256, 256, True, True),       # configs[1]-sized quantizer behind a 256-channel encoder
    (3, 37, 6, 8, True, True),             # the golden VQVAE's sizes: nothing is a multiple of a tile
    (5, 129, 100, 72, True, False),
    (1, 700, 192, 300, False, True),       # rows layout (the codebook side of the decoder projection), D > 256
    (2, 50, 520, 64, True, True),          # C > 512: 27 k-blocks of split operands stream through the ring
])
def test_proj_forward_vs_fp64(cuda_dev, n_lines, frames, C, D, cf, bias):
    from pero_pretraining_b200 import ops
    g = torch.Generator().manual_seed(n_lines * 1000 + C)
    N = n_lines * frames
    x = torch.randn(n_lines, C, frames, generator=g) if cf else torch.randn(N, C, generator=g)
    w = torch.randn(D, C, generator=g) / np.sqrt(C)
    b = torch.randn(D, generator=g) if bias else None
    packed = torch.zeros(N, dtype=torch.int64, device=cuda_dev)
    rows, xb = ops.proj_forward(x.to(cuda_dev), w.to(cuda_dev), None if b is None else b.to(cuda_dev), n_lines, frames, cf,
                                want_rows=True, want_bf16=True, packed=packed)
    torch.cuda.synchronize()
    xr = (x.permute(0, 2, 1).reshape(N, C) if cf else x).double()
    ref = xr @ w.double().t() + (b.double() if bias else 0.0)
    bound = 4e-5 * (xr.abs() @ w.double().abs().t()) + 1e-7
    err = (rows.cpu().double() - ref).abs()
    assert bool((err <= bound).all()), f"max err {err.max():.3e}, worst ratio {(err / bound).max():.2f}"
    Dp = (D + 63) // 64 * 64
    assert xb.shape == (N, Dp)
    assert torch.equal(xb[:, :D].cpu(), rows.cpu().bfloat16()), "the bf16 operand is the rounded fp32 row"
    assert not bool(xb[:, D:].any()), "padding columns of the operand must be zero"
    assert bool((packed == INT64_MAX).all()), "packed winners reset to empty"
    # each output alone
    rows2, none = ops.proj_forward(x.to(cuda_dev), w.to(cuda_dev), None if b is None else b.to(cuda_dev), n_lines, frames, cf)
    assert none is None and torch.equal(rows2, rows)
    none, xb2 = ops.proj_forward(x.to(cuda_dev), w.to(cuda_dev), None if b is None else b.to(cuda_dev), n_lines, frames, cf,
                                 want_rows=False, want_bf16=True)
    assert none is None and torch.equal(xb2, xb)


def test_gather_rows_cf_is_exact(cuda_dev):
    from pero_pretraining_b200 import ops
    g = torch.Generator().manual_seed(5)
    for n_lines, frames, K, C in ((3, 37, 50, 6), (2, 129, 1000, 200), (64, 128, 8192, 256)):
        table = torch.randn(K, C, generator=g)
        idx = torch.randint(0, K, (n_lines * frames,), generator=g)
        out = ops.gather_rows_cf(table.to(cuda_dev), idx.to(cuda_dev), n_lines, frames)
        expect = table[idx].view(n_lines, frames, C).permute(0, 2, 1)
        assert torch.equal(out.cpu(), expect)


class _Enc(torch.nn.Module):
    def __init__(self, c):
        super().__init__()
        self.out_channels = c

    def forward(self, x):
        return x


class _Dec(torch.nn.Module):
    def __init__(self, c):
        super().__init__()
        self.base_channels = c

    def forward(self, x):
        return x


@pytest.mark.parametrize("n_lines,H,W,C,K,D,decay", [
    (4, 2, 33, 24, 64, 16, 0.99),
    (64, 1, 128, 256, 8192, 256, 0.99),      # configs[1] behind 256-channel projections
    (8, 1, 128, 96, 512, 64, 0.0),           # no EMA: the codebook is trained by the loss instead
])
def test_fused_quantize_matches_conv_modules(cuda_dev, n_lines, H, W, C, K, D, decay):
    """VQVAE.quantize fused (projection GEMM -> distance GEMM -> projected-codebook gather) against the same module with
    fuse_projections = False (torch.nn.Conv2d around VectorQuantizer, TF32 off): labels (near-tie rule on the projected
    features), tokens, EMA state, and the gradients of both projections and of the features."""
    import copy
    from pero_pretraining_b200 import VQVAE
    torch.manual_seed(n_lines + K)
    a = VQVAE(_Enc(C), _Dec(C), K, D, 0.25, decay).to(cuda_dev).train()
    N = n_lines * H * W
    with torch.no_grad():
        a.encoder_projection_layer.weight.normal_(0, 1.0 / np.sqrt(C))
        if decay == 0.0:
            a.vq.embedding.weight.normal_()
        # features whose PROJECTION lies near a codeword (codeword + 0.3 * noise, pulled back through the pseudo-inverse
        # of the projection), so that the two arithmetic paths cannot disagree on a label by a near-tie
        we = a.encoder_projection_layer.weight.view(D, C).double()
        target = a.vq.embedding.weight[torch.randint(0, K, (N,), device=cuda_dev)].double() + 0.3 * torch.randn(N, D, device=cuda_dev).double()
        rows = (target - a.encoder_projection_layer.bias.double()) @ torch.linalg.pinv(we).t()
        feats = rows.float().view(n_lines, H, W, C).permute(0, 3, 1, 2).contiguous()
    b = copy.deepcopy(a)
    a.fuse_projections, b.fuse_projections = True, False
    g_out = torch.randn(n_lines, C, H, W, device=cuda_dev)
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        outs = []
        for m in (a, b):
            f = feats.clone().requires_grad_(True)
            tokens, labels = m.quantize(f)
            loss = m.vq.calculate_loss(tokens, f) + (tokens * g_out).sum()
            loss.backward()
            outs.append((tokens.detach(), labels, f.grad, loss.detach()))
        # near-tie rule on the projected features (fp64 brute force on the device)
        with torch.no_grad():
            xp = torch.nn.functional.conv2d(feats.double(), b.encoder_projection_layer.weight.double(),
                                            b.encoder_projection_layer.bias.double()).permute(0, 2, 3, 1).reshape(-1, D)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    (ta, la, ga, lossa), (tb, lb, gb, lossb) = outs
    assert ta.shape == tb.shape == (n_lines, C, H, W) and la.dtype == torch.int64 and la.shape == lb.shape
    differs = la != lb
    flip = float(differs.float().mean())
    assert flip <= 0.002, f"{flip:.4f} of the labels differ"
    same = ~differs
    sel = same.view(n_lines, 1, H, W).expand_as(ta)
    scale = float(tb.abs().max())
    assert float((ta - tb)[sel].abs().max()) <= 1e-4 * scale, "decoder projection of the quantized frames"
    if not bool(differs.any()):
        np.testing.assert_allclose(lossa.item(), lossb.item(), rtol=1e-4)
        for name in ("encoder_projection_layer.weight", "encoder_projection_layer.bias", "decoder_projection_layer.weight",
                     "decoder_projection_layer.bias"):
            pa, pb = a.get_parameter(name).grad, b.get_parameter(name).grad
            assert float((pa - pb).abs().max()) <= 1e-3 * float(pb.abs().max()) + 1e-6, name
        assert float((ga - gb).abs().max()) <= 1e-3 * float(gb.abs().max()) + 1e-6, "gradient of the features"
        # (a cold-start EMA step divides by cluster sizes near epsilon: codewords reach 1e2, hence the absolute term)
        wa, wb = a.vq.embedding.weight.detach(), b.vq.embedding.weight.detach()
        assert float((wa - wb).abs().max()) <= 1e-4 * float(wb.abs().max()), "codebook after the EMA update"
    # eval ('auto' takes the fused path when no gradient is wanted): labels only (label production), no EMA update, and the
    # decoder-projected codebook is computed once and kept until the codebook or the projection changes
    a.fuse_projections = 'auto'
    a.eval(); b.eval()
    b.load_state_dict(a.state_dict())
    w0 = a.vq.embedding.weight.detach().clone()
    with torch.no_grad():
        t2, l2 = a.quantize(feats)
        table = a._table
        assert table is not None and tuple(table.shape) == (K, C)
        t3, l3 = a.quantize(feats)
        assert a._table is table and torch.equal(t3, t2) and torch.equal(l3, l2)
        torch.backends.cudnn.allow_tf32 = False                # the comparison path's convolutions in fp32
        try:
            tb2, lb2 = b.quantize(feats)
        finally:
            torch.backends.cudnn.allow_tf32 = prev[0]
        a.decoder_projection_layer.bias.add_(1.0)             # in-place update bumps the version: the table is rebuilt
        t4, _ = a.quantize(feats)
        assert a._table is not table
    assert torch.equal(a.vq.embedding.weight.detach(), w0)
    assert l2.shape == la.shape and t2.shape == ta.shape
    same = (l2 == lb2)
    assert float(same.float().mean()) >= 0.998
    sel = same.view(n_lines, 1, H, W).expand_as(t2)
    assert float((t2 - tb2)[sel].abs().max()) <= 1e-4 * float(tb2.abs().max())
    assert float((t4 - t2 - 1.0).abs().max()) <= 1e-5 * (1.0 + float(t2.abs().max()))
    # with gradients wanted, 'auto' keeps the torch.nn.Conv2d projections (cuDNN) around the quantizer
    a.train()
    f = feats.clone().requires_grad_(True)
    tok, _ = a.quantize(f)
    assert tok.grad_fn is not None and "ProjectedQuantize" not in type(tok.grad_fn).__name__


def test_proj_entry_points_stay_inside_their_buffers(cuda_dev):
    from pero_pretraining_b200 import _lib
    from test_gpu_canaries import Guarded, _s
    L = _lib.lib()
    gen = torch.Generator().manual_seed(11)
    for n_lines, frames, C, D, cf in ((3, 37, 6, 8, 1), (2, 129, 100, 72, 1), (1, 300, 70, 260, 0)):
        N = n_lines * frames
        g = Guarded(cuda_dev)
        x = (torch.randn(n_lines, C, frames, generator=gen) if cf else torch.randn(N, C, generator=gen)).to(cuda_dev)
        w = torch.randn(D, C, generator=gen).to(cuda_dev)
        b = torch.randn(D, generator=gen).to(cuda_dev)
        rows = g.buf(N * D * 4, torch.float32)
        Dp = (D + 63) // 64 * 64
        xb = g.buf(N * Dp * 2, torch.bfloat16)
        packed = g.buf(N * 8, torch.int64)
        wsb = L.pero_proj_workspace_bytes(N, C, D)
        ws = g.buf(wsb)
        _lib.check(L.pero_proj_forward(x.data_ptr(), n_lines, frames, cf, C, w.data_ptr(), b.data_ptr(), D, rows.data_ptr(),
                                       xb.data_ptr(), packed.data_ptr(), ws.data_ptr(), wsb, _s()), "proj")
        idx = torch.randint(0, N, (N,), generator=gen).to(cuda_dev)
        out = g.buf(n_lines * D * frames * 4, torch.float32)
        _lib.check(L.pero_gather_rows_cf(rows.data_ptr(), idx.data_ptr(), n_lines, frames, N, D, out.data_ptr(), _s()), "gather")
        g.check(f"projection entry points {n_lines}x{frames} C={C} D={D}")
        assert L.pero_proj_forward(x.data_ptr(), n_lines, frames, cf, C, w.data_ptr(), b.data_ptr(), D, rows.data_ptr(),
                                   xb.data_ptr(), packed.data_ptr(), ws.data_ptr(), wsb - 1, _s()) == -3      # PERO_ERR_WORKSPACE


def test_label_production_matches_quantize(cuda_dev):
    """compute_labels (scripts/produce_vqvae_labels.py:25-44) through VQVAE.labels -- encoder projection + assignment only --
    gives the labels VQVAE.quantize gives, filtered by the image masks."""
    from pero_pretraining_b200 import VQVAE, compute_labels
    torch.manual_seed(3)
    m = VQVAE(_Enc(24), _Dec(24), 64, 16, 0.25, 0.99).to(cuda_dev).eval()
    batches = []
    for b in range(3):
        images = torch.randn(4, 24, 1, 37, device=cuda_dev)
        masks = [(np.arange(37) < 30 + i).astype(int) for i in range(4)]
        batches.append({"images": images, "ids": [f"line{b}_{i}" for i in range(4)], "image_masks": masks})
    data = compute_labels(m, batches)
    assert len(data) == 12
    with torch.no_grad():
        for batch in batches:
            _, ref = m.quantize(batch["images"])
            ref = ref.reshape(4, 37).cpu().numpy()
            for i, line_id in enumerate(batch["ids"]):
                assert data[line_id] == ref[i][batch["image_masks"][i] == 1].tolist()
    w0 = m.vq.embedding.weight.detach().clone()
    m.train()
    lab = m.labels(batches[0]["images"])                        # never updates the EMA state, also in training mode
    assert torch.equal(m.vq.embedding.weight.detach(), w0) and lab.dtype == torch.int64 and lab.shape == (4 * 37,)


@pytest.mark.parametrize("fuse", [True, "auto", False])
def test_vqvae_quantize_eval_golden(cuda_dev, fuse):
    """The reference's own VQVAE.quantize outputs (eval mode; tests/golden/vqvae_quantize_eval.npz, generated by importing
    the reference) against the drop-in module: labels exactly (every frame of the fixture has a clear fp64 gap), projected
    tokens to 1e-4 of their range -- through the fused projections (True / 'auto') and through the Conv2d path (False)."""
    from pero_pretraining_b200 import VQVAE
    g = load_golden("vqvae_quantize_eval")
    C, D, K, nl, H, W = [int(v) for v in g["dims"]]
    m = VQVAE(_Enc(C), _Dec(C), K, D)
    m.load_state_dict({k[len("state_"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("state_")})
    m = m.to(cuda_dev).eval()
    m.fuse_projections = fuse
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            tokens, labels = m.quantize(torch.from_numpy(g["features"]).to(cuda_dev))
            only_labels = m.labels(torch.from_numpy(g["features"]).to(cuda_dev))
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    assert float(g["gap"].min()) > EPS_TIE
    assert np.array_equal(labels.cpu().numpy(), g["out_labels"])
    assert np.array_equal(only_labels.cpu().numpy(), g["out_labels"])
    ref = g["out_tokens"]
    assert tokens.shape == ref.shape
    assert float(np.abs(tokens.cpu().numpy() - ref).max()) <= 1e-4 * float(np.abs(ref).max())
