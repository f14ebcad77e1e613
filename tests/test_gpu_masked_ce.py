"""GPU parity of the fused LinearHead + MaskedCrossEntropyLoss path (and the logits-in loss) against the
golden fixture produced by the reference and against the oracle.

Stated tolerances (north_star: "losses and gradients agree within a stated bf16/tf32 relative tolerance"):
the fused path multiplies bf16-rounded hidden states and weights with fp32 accumulation, and its backward
rounds dlogits to bf16 before the d_W / d_h GEMMs, so
    loss:       |got - ref| <= 3e-3 * |ref|
    gradients:  max|got - ref| <= 2e-2 * max|ref|   (per tensor)
The logits-in loss is fp32 end to end: rtol 1e-5."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import pero_oracle as O

pytestmark = pytest.mark.gpu
LOSS_RTOL = 3e-3
GRAD_RTOL = 2e-2


def T(a):
    return torch.from_numpy(np.asarray(a))


def _close_grad(got, ref, what):
    got, ref = got.detach().double().cpu(), ref.double()
    err = (got - ref).abs().max().item()
    assert err <= GRAD_RTOL * max(ref.abs().max().item(), 1e-12), f"{what}: max err {err:.3e} vs max|ref| {ref.abs().max().item():.3e}"


def _run_fused(dev, h, W, b, labels, mask, uw, h_dtype=torch.float32):
    from pero_pretraining_b200 import LinearHead
    head = LinearHead(W.shape[1], W.shape[0]).to(dev)
    with torch.no_grad():
        head.linear.weight.copy_(W)
        head.linear.bias.copy_(b)
    hh = h.to(dev).to(h_dtype).requires_grad_(True)
    loss = head.masked_loss(hh, labels.to(dev), mask, uw)
    loss.backward()
    torch.cuda.synchronize()
    return loss, hh.grad, head.linear.weight.grad, head.linear.bias.grad


def test_fused_head_ce_golden(cuda_dev):
    g = load_golden("masked_ce")
    h, W, b, labels = T(g["h"]), T(g["W"]), T(g["b"]), T(g["labels"])
    for tag, uw in (("plain", None), ("unmasked", float(g["unmasked_weight"]))):
        for mask in (g["mask"], T(g["mask"])):        # numpy mask (host compaction) and tensor mask (device compaction)
            loss, gh, gW, gb = _run_fused(cuda_dev, h, W, b, labels, mask, uw)
            assert abs(loss.item() - float(g[f"{tag}_loss"])) <= LOSS_RTOL * abs(float(g[f"{tag}_loss"]))
            _close_grad(gh, T(g[f"{tag}_g_h"]), "d_h")
            _close_grad(gW, T(g[f"{tag}_g_W"]), "d_W")
            _close_grad(gb, T(g[f"{tag}_g_b"]), "d_b")
            pad = (T(g["labels"]) < 0).to(gh.device)
            assert float(gh[pad].abs().max()) == 0.0            # padding frames get no gradient


@pytest.mark.parametrize("Nl,T_,Dh,V,p,dtype", [
    (8, 128, 512, 4096, 0.15, torch.float32),      # config c1
    (8, 128, 512, 4096, 0.15, torch.bfloat16),     # --bfloat16 autocast hands the head bf16 hidden states
    (32, 128, 512, 4096, 0.15, torch.bfloat16),    # config c3: one rank's share (32 lines) of the 256-line batch
    (64, 128, 512, 8192, 0.15, torch.float32),     # the shape every bench.py number is quoted on (configs[1] labels)
    (64, 128, 512, 8192, 0.15, torch.bfloat16),
    (5, 37, 96, 1000, 0.5, torch.float32),         # ragged: V, Dh, M not multiples of any tile
    (2, 9, 64, 300, 1.0, torch.float32),           # every frame masked
    (3, 50, 512, 257, 0.02, torch.float32),        # a handful of masked frames
    (4, 64, 768, 1000, 0.3, torch.float32),        # model_dim 768: both operands stream through the ring (no resident A)
    (2, 64, 1024, 520, 0.5, torch.bfloat16),       # model_dim 1024
])
def test_fused_head_ce_vs_oracle(cuda_dev, Nl, T_, Dh, V, p, dtype):
    rng = np.random.default_rng(Nl * 1000 + V)
    g = torch.Generator(device="cpu").manual_seed(V + Dh)
    h = torch.randn(Nl, T_, Dh, generator=g)
    bound = 1.0 / np.sqrt(Dh)
    W = (torch.rand(V, Dh, generator=g) * 2 - 1) * bound
    b = (torch.rand(V, generator=g) * 2 - 1) * bound
    labels = torch.from_numpy(rng.integers(0, V, size=(Nl, T_))).long()
    if T_ > 8:
        labels[:, -4:] = -1
    mask = O.create_mask(labels.numpy(), p, rng)
    if mask.sum() == 0:
        mask[0, 0] = 1
    if dtype == torch.bfloat16:
        h = h.bfloat16().float()           # the oracle sees the same (bf16-representable) hidden states
    ref_loss, ref_dh, ref_dW, ref_db = O.head_masked_ce(h, W, b, labels, T(mask))
    loss, gh, gW, gb = _run_fused(cuda_dev, h, W, b, labels, mask, None, dtype)
    assert abs(loss.item() - ref_loss.item()) <= LOSS_RTOL * abs(ref_loss.item())
    assert gh.dtype == dtype
    _close_grad(gh.float(), ref_dh, "d_h")
    _close_grad(gW, ref_dW, "d_W")
    _close_grad(gb, ref_db, "d_b")


def test_fused_head_ce_is_deterministic(cuda_dev):
    g = torch.Generator(device="cpu").manual_seed(77)
    h, W, b = torch.randn(4, 64, 512, generator=g), torch.randn(2048, 512, generator=g) * 0.04, torch.zeros(2048)
    labels = torch.randint(0, 2048, (4, 64), generator=g)
    mask = (torch.rand(4, 64, generator=g) < 0.3).long().numpy()
    a = _run_fused(cuda_dev, h, W, b, labels, mask, None)
    c = _run_fused(cuda_dev, h, W, b, labels, mask, None)
    for u, v in zip(a, c):
        assert torch.equal(u, v)


def test_empty_mask_gives_nan_like_reference(cuda_dev):
    g = torch.Generator(device="cpu").manual_seed(1)
    h, W, b = torch.randn(2, 8, 64, generator=g), torch.randn(100, 64, generator=g), torch.zeros(100)
    labels = torch.randint(0, 100, (2, 8), generator=g)
    loss, gh, gW, gb = _run_fused(cuda_dev, h, W, b, labels, np.zeros((2, 8), dtype=int), None)
    assert torch.isnan(loss)


def test_logits_in_loss_golden(cuda_dev):
    from pero_pretraining_b200 import MaskedCrossEntropyLoss
    g = load_golden("masked_ce")
    for tag, uw in (("plain", None), ("unmasked", float(g["unmasked_weight"]))):
        logits = T(g[f"{tag}_logits"]).to(cuda_dev).requires_grad_(True)
        loss = MaskedCrossEntropyLoss(uw)(logits, T(g["labels"]).to(cuda_dev), T(g["mask"]).to(cuda_dev))
        loss.backward()
        np.testing.assert_allclose(loss.item(), float(g[f"{tag}_loss"]), rtol=1e-5)
        np.testing.assert_allclose(logits.grad.cpu().numpy(), g[f"{tag}_g_logits"], rtol=1e-4, atol=1e-7)


def test_model_glue_train_and_eval(cuda_dev):
    """MaskedTransformerEncoder.forward contract (model.py:41-56): dict keys; loss in training without
    materialising logits; logits for every frame in eval; numpy or tensor mask."""
    from pero_pretraining_b200 import LinearHead, MaskedTransformerEncoder

    class Backbone(torch.nn.Module):          # stand-in for the (out-of-scope) transformer: [n, 3, 40, W] -> [n, 64, W/8]
        def __init__(self):
            super().__init__()
            self.conv = torch.nn.Conv2d(3, 64, (40, 8), stride=(40, 8))

        def forward(self, images, mask=None):
            return self.conv(images).squeeze(2)

    torch.manual_seed(0)
    model = MaskedTransformerEncoder(Backbone(), LinearHead(64, 200)).to(cuda_dev)
    images = torch.rand(4, 3, 40, 256, device=cuda_dev)
    labels = torch.randint(0, 200, (4, 32), device=cuda_dev)
    mask = (np.random.default_rng(0).random((4, 32)) < 0.3).astype(int)
    model.train()
    out = model(images, labels, mask)
    assert set(out.keys()) == {"output", "loss"} and out["output"] is None and out["loss"].dim() == 0
    out["loss"].backward()
    assert model.backbone.conv.weight.grad is not None and model.head.linear.weight.grad is not None
    model.eval()
    with torch.no_grad():
        ev = model(images, labels, torch.from_numpy(mask).to(cuda_dev))
    assert ev["output"].shape == (4, 32, 200)
    ref = O.masked_ce(ev["output"].float().cpu(), labels.cpu(), T(mask))
    assert abs(ev["loss"].item() - ref.item()) <= LOSS_RTOL * abs(ref.item())
    assert abs(out["loss"].item() - ref.item()) <= LOSS_RTOL * abs(ref.item())
    assert model(images)["loss"] is None
    # predictions of every frame without the logits (visualizer.py:32)
    pred = model.predict(images)
    assert pred.shape == (4, 32) and (pred == ev["output"].argmax(-1)).float().mean().item() > 0.97


def test_model_glue_device_pixel_masking(cuda_dev):
    """MaskedTransformerEncoder(pixel_masker=PixelMasker()): the backbone sees exactly the pixels the reference's
    backbone.mask() would produce (models/transformers.py:45-68), from ONE staged masked-frame list shared with the
    loss; same loss as the path in which the backbone masks its own input."""
    from pero_pretraining_b200 import LinearHead, MaskedTransformerEncoder, PixelMasker

    class Backbone(torch.nn.Module):          # masks its own input like the reference's TransformerEncoder.forward
        def __init__(self):
            super().__init__()
            self.conv = torch.nn.Conv2d(3, 64, (40, 8), stride=(40, 8))
            self.tile = T(O.mask_tile(3, (40, 8)))
            self.seen = None

        def forward(self, images, mask=None):
            if mask is not None:
                images = T(O.mask_pixels(images.cpu().numpy(), np.asarray(mask), self.tile.numpy())).to(images.device)
            self.seen = images.detach().clone()
            return self.conv(images).squeeze(2)

    torch.manual_seed(0)
    head = LinearHead(64, 200)
    a = MaskedTransformerEncoder(Backbone(), head).to(cuda_dev).train()
    b = MaskedTransformerEncoder(a.backbone, head, pixel_masker=PixelMasker()).to(cuda_dev).train()
    images = torch.rand(4, 3, 40, 256, device=cuda_dev)
    labels = torch.randint(0, 200, (4, 32), device=cuda_dev)
    mask = (np.random.default_rng(1).random((4, 32)) < 0.3).astype(int)
    la = a(images.clone(), labels, mask)["loss"]
    seen_a = a.backbone.seen
    lb = b(images.clone(), labels, mask)["loss"]
    assert torch.equal(seen_a, b.backbone.seen)
    assert torch.equal(la, lb)


def test_mask_compact_matches_nonzero(cuda_dev):
    from pero_pretraining_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(4)
    for dtype in (torch.int64, torch.int32, torch.uint8, torch.bool):
        m = (torch.rand(5, 777, generator=g) < 0.2).to(dtype).to(cuda_dev)
        labels = torch.randint(-1, 5, (5, 777), generator=g).to(cuda_dev)
        rows, count = ops.mask_compact(m, None, 1)
        ref = torch.nonzero(m.reshape(-1).long() == 1).flatten()
        assert int(count) == ref.numel() and torch.equal(rows[:ref.numel()].long(), ref)
        rows0, count0 = ops.mask_compact(m, labels, 0)
        ref0 = torch.nonzero((m.reshape(-1).long() == 0) & (labels.reshape(-1) >= 0)).flatten()
        assert int(count0) == ref0.numel() and torch.equal(rows0[:ref0.numel()].long(), ref0)


@pytest.mark.parametrize("keep", [False, True])
def test_backward_by_label_ranges_equals_full_backward(cuda_dev, keep):
    """pero_masked_ce_bwd_range: walking the label axis range by range (what the data-parallel step does to overlap
    the exchange of d_W with the next range) gives bit-identical d_W and d_b, and d_h to fp32 summation order -- with the logits GEMM recomputed
    (keep=False) and with the forward's bf16 logits converted in place (keep=True, PERO_CE_KEEP_LOGITS)."""
    from pero_pretraining_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(5)
    N, Dh, V, M = 1024, 512, 1300, 150
    h = torch.randn(N, Dh, generator=g).to(cuda_dev)
    W = (torch.randn(V, Dh, generator=g) * 0.05).to(cuda_dev)
    b = (torch.randn(V, generator=g) * 0.1).to(cuda_dev)
    labels = torch.randint(0, V, (N,), generator=g).to(cuda_dev)
    rows = torch.sort(torch.randperm(N, generator=g)[:M]).values.int().to(cuda_dev)
    head = ops.PreparedHead(V, Dh, cuda_dev).prepare(W, b)
    loss_sum, lse, ws = ops.masked_ce_fwd(h, rows, labels, head, keep_logits=keep)
    d_h, d_W, d_b = ops.masked_ce_bwd(h, rows, labels, head, lse, None, 1.0 / M, ws=ws, ws_from_fwd=True, logits_in_ws=keep)
    loss_sum2, lse2, ws2 = ops.masked_ce_fwd(h, rows, labels, head, keep_logits=keep)
    assert torch.equal(lse, lse2) and torch.equal(loss_sum, loss_sum2)
    flat = torch.full((V * Dh + V,), float("nan"), device=cuda_dev)
    for v0, v1 in ((0, 512), (512, 1024), (1024, V)):
        ops.masked_ce_bwd(h, rows, labels, head, lse2, None, 1.0 / M, ws=ws2, ws_from_fwd=True, want_dh=False, flat_out=flat,
                          return_flat=True, v_range=(v0, v1), logits_in_ws=keep)
    d_h2, _, _ = ops.masked_ce_bwd(h, rows, labels, head, lse2, None, 1.0 / M, ws=ws2, want_dw=False)
    assert torch.equal(flat[:V * Dh].view(V, Dh), d_W)
    assert torch.equal(flat[V * Dh:], d_b)
    # d_h = dlogits @ W is summed over label-axis slices whose number depends on how many SM pairs the GEMM has to itself
    # (alone here, beside the d_W GEMM in the one-call backward): same dlogits bits, fp32 summation order differs
    assert torch.equal(d_h2 == 0, d_h == 0)
    assert float((d_h2 - d_h).abs().max()) <= 2e-6 * float(d_h.abs().max())
    d_h3, _, _ = ops.masked_ce_bwd(h, rows, labels, head, lse2, None, 1.0 / M, ws=ws2, want_dw=False)
    assert torch.equal(d_h3, d_h2)                    # and bit-identical from run to run
    with pytest.raises(Exception):                    # ranges must start on a multiple of 256
        ops.masked_ce_bwd(h, rows, labels, head, lse2, None, 1.0 / M, ws=ws2, want_dh=False, v_range=(100, 512))


@pytest.mark.parametrize("N,Dh,V,M,dtype", [(8192, 512, 8192, 1245, torch.float32), (1024, 512, 4096, 150, torch.bfloat16),
                                             (300, 96, 1000, 77, torch.float32), (64, 576, 300, 64, torch.float32)])
def test_kept_logits_backward_matches_recompute_backward(cuda_dev, N, Dh, V, M, dtype):
    """The two backward routes of the fused head -- logits GEMM recomputed, or the forward's kept softmax numerators
    (bf16, relative to per-chunk maxima) scaled in place -- agree within the bf16 tolerance of the path (both round the
    dlogits to bf16 once more or less), give the same loss terms, and the in-place route's d_b sums to zero."""
    from pero_pretraining_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(N + V)
    h = torch.randn(N, Dh, generator=g).to(cuda_dev).to(dtype)
    W = (torch.randn(V, Dh, generator=g) * 0.08).to(cuda_dev)        # logits up to ~ +-8
    b = (torch.randn(V, generator=g) * 0.1).to(cuda_dev)
    labels = torch.randint(0, V, (N,), generator=g).to(cuda_dev)
    rows = torch.sort(torch.randperm(N, generator=g)[:M]).values.int().to(cuda_dev)
    head = ops.PreparedHead(V, Dh, cuda_dev).prepare(W, b)
    out = {}
    for keep in (False, True):
        loss_sum, lse, ws = ops.masked_ce_fwd(h, rows, labels, head, keep_logits=keep)
        out[keep] = (loss_sum.clone(), lse.clone()) + tuple(
            t.float().clone() for t in ops.masked_ce_bwd(h, rows, labels, head, lse, None, 1.0 / M, ws=ws, ws_from_fwd=True,
                                                         logits_in_ws=keep))
    # the loss comes from the fp32 accumulators on both routes (summed chunk by chunk on the kept route: last-bit differences)
    assert torch.allclose(out[False][0], out[True][0], rtol=1e-5) and torch.allclose(out[False][1], out[True][1], rtol=1e-5, atol=1e-5)
    for i, what in ((2, "d_h"), (3, "d_W"), (4, "d_b")):
        a, c = out[False][i].double(), out[True][i].double()
        err = (a - c).abs().max().item()
        assert err <= GRAD_RTOL * max(a.abs().max().item(), 1e-12), f"{what}: {err:.3e} vs {a.abs().max().item():.3e}"
    # rows that are not masked get exactly zero gradient on both routes
    unmasked = torch.ones(N, dtype=torch.bool, device=cuda_dev)
    unmasked[rows.long()] = False
    assert float(out[True][2][unmasked].abs().max()) == 0.0 if bool(unmasked.any()) else True
    # gradient of the bias = sum over the masked frames of (softmax - onehot) / M: the columns sum to zero overall
    assert abs(float(out[True][4].double().sum())) < 1e-3


@pytest.mark.parametrize("Nl,T_,Dh,V,p", [(8, 128, 512, 4096, 0.15), (5, 37, 96, 1000, 0.5), (3, 50, 512, 257, 0.3)])
def test_masked_errors_match_tester(cuda_dev, Nl, T_, Dh, V, p):
    """Fused evaluation (label rank counted in the logits GEMM's epilogue) vs Tester._update_errors restated in the
    oracle on full logits (masked_pretraining/tester.py:70-93).  A frame may differ only where the label's logit is
    within bf16 noise of the k-th largest logit."""
    from pero_pretraining_b200 import LinearHead, ops
    from pero_pretraining_b200.masked_pretraining import update_errors
    rng = np.random.default_rng(V)
    g = torch.Generator(device="cpu").manual_seed(V + 3)
    h = torch.randn(Nl, T_, Dh, generator=g)
    W = torch.randn(V, Dh, generator=g) * (2.0 / np.sqrt(Dh))
    b = torch.randn(V, generator=g) * 0.1
    labels = torch.from_numpy(rng.integers(0, V, size=(Nl, T_))).long()
    # make the task non-trivial: every frame is pulled towards its label's weight row by a random amount, so the
    # label's rank ranges from 0 to the hundreds
    pull = torch.from_numpy(rng.random((Nl, T_))).float()[..., None] * 4.5
    h = h + pull * torch.nn.functional.normalize(W[labels], dim=-1)
    mask = (rng.random((Nl, T_)) < p).astype(int)
    head = LinearHead(Dh, V).to(cuda_dev)
    with torch.no_grad():
        head.linear.weight.copy_(W); head.linear.bias.copy_(b)
    res = head.masked_errors(h.to(cuda_dev), labels.to(cuda_dev), mask, ks=(1, 3, 10))
    logits = O.linear_head(h.bfloat16().float(), W.bfloat16().float(), b)         # same operand rounding as the device
    ref = O.topk_errors(logits.numpy(), labels.numpy(), mask, ks=(1, 3, 10))
    assert res["length"] == ref["length"]
    # frames whose label logit is within 1e-3 (relative to the logit scale) of the k-th largest may flip
    z = logits.numpy()[mask == 1]; y = labels.numpy()[mask == 1]
    zl = z[np.arange(len(y)), y]
    srt = -np.sort(-z, axis=1)
    for k in (1, 3, 10):
        near = int((np.abs(zl - srt[:, k - 1]) < 1e-3 * np.abs(srt[:, 0])).sum()) + int((np.abs(zl - srt[:, k]) < 1e-3 * np.abs(srt[:, 0])).sum())
        assert abs(int(res[f"errors_{k}"]) - ref[f"errors_{k}"]) <= near, (k, int(res[f"errors_{k}"]), ref[f"errors_{k}"], near)
    ref_loss = O.masked_ce(logits, labels, T(mask))
    assert abs(res["loss"].item() - ref_loss.item()) <= LOSS_RTOL * abs(ref_loss.item())
    assert 0 < ref["errors_10"] <= ref["errors_3"] <= ref["errors_1"] < ref["length"]
    if V >= 1000:
        assert ref["errors_10"] < ref["errors_1"]                                   # the case separates the three counts
    # rank itself: exact where the logit gaps are clear
    rows = torch.from_numpy(np.flatnonzero(mask.reshape(-1) == 1).astype(np.int32)).to(cuda_dev)
    _, _, rank, _ = ops.masked_ce_eval(h.to(cuda_dev).reshape(-1, Dh), rows, labels.to(cuda_dev).reshape(-1), head._prepared(),
                                       ks=(1,), want_rank=True)
    ref_rank = (z > zl[:, None]).sum(1)
    margin = np.abs(z - zl[:, None]); margin[np.arange(len(y)), y] = np.inf
    clear = margin.min(1) > 1e-3 * np.abs(srt[:, 0])
    assert np.array_equal(rank.cpu().numpy()[clear], ref_rank[clear])
    acc = update_errors(update_errors({}, res), res)
    assert acc["length"] == 2 * res["length"] and int(acc["errors_3"]) == 2 * int(res["errors_3"])


def test_pixel_masking_golden_and_shared_rows(cuda_dev):
    """PixelMasker (TransformerEncoder.mask, models/transformers.py:53-68) bit-exact against the reference's output, from a
    numpy mask, a tensor mask and a pre-staged row list; ragged width (W not a multiple of 8) against the oracle."""
    from pero_pretraining_b200 import PixelMasker, masked_rows
    g = load_golden("pixel_mask")
    pm = PixelMasker().to(cuda_dev)
    assert torch.equal(pm.mask_tile.cpu(), T(g["tile"]))
    for mask in (g["mask"], T(g["mask"]).to(cuda_dev)):
        x = T(g["x"]).to(cuda_dev)
        out = pm(x, mask)
        assert out.data_ptr() == x.data_ptr()                   # in place, like the reference
        assert torch.equal(out.cpu(), T(g["masked"]))
    x = T(g["x"]).to(cuda_dev)
    assert torch.equal(pm(x, rows=masked_rows(g["mask"], cuda_dev)).cpu(), T(g["masked"]))
    x = T(g["x"]).to(cuda_dev)
    assert torch.equal(pm(x, np.zeros_like(g["mask"])).cpu(), T(g["x"]))       # empty mask: untouched
    rng = np.random.default_rng(3)
    xr = rng.random((2, 3, 40, 77), dtype=np.float32)
    mr = (rng.random((2, 10)) < 0.5).astype(int)
    mr[1, 9] = 1                                                # the last, partial column
    got = pm(T(xr).to(cuda_dev), mr).cpu().numpy()
    np.testing.assert_array_equal(got, O.mask_pixels(xr, mr, g["tile"]))


def test_head_argmax_without_logits(cuda_dev):
    """LinearHead.argmax == torch.argmax(head(hidden), -1) (masked_pretraining/visualizer.py:32) except where the two best
    logits are within bf16 rounding; exact ties -> lowest label; the golden logits reproduce exactly."""
    from pero_pretraining_b200 import LinearHead
    g = torch.Generator(device="cpu").manual_seed(5)
    for Nl, T_, Dh, V in [(4, 50, 512, 4096), (3, 17, 96, 1000), (2, 128, 768, 300)]:
        head = LinearHead(Dh, V).to(cuda_dev)
        h = torch.randn(Nl, T_, Dh, generator=g).to(cuda_dev)
        got = head.argmax(h)
        z = head(h).double()
        ref = torch.argmax(z, dim=-1)
        assert got.shape == ref.shape and got.dtype == torch.int64
        top2 = torch.topk(z, 2, dim=-1).values
        gap = (top2[..., 0] - top2[..., 1]) / top2[..., 0].abs().clamp_min(1e-6)
        differs = got != ref
        assert float(differs.float().mean()) < 0.02
        assert not bool(differs.any()) or float(gap[differs].max()) < 2e-2
    # one-hot-like head: logits are exact in bf16, duplicates resolve to the lowest label
    head = LinearHead(64, 128).to(cuda_dev)
    with torch.no_grad():
        head.linear.weight.zero_(); head.linear.bias.zero_()
        head.linear.weight[7, 3] = 2.0; head.linear.weight[31, 3] = 2.0; head.linear.weight[90, 5] = 4.0
    h = torch.zeros(1, 3, 64, device=cuda_dev)
    h[0, 0, 3] = 1.0; h[0, 1, 5] = 1.0
    assert head.argmax(h).tolist() == [[7, 90, 0]]
