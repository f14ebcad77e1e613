"""Pins oracle/pero_oracle.py against the reference's own outputs (tests/golden/*.npz, produced by
tests/golden/make_golden.py executing /root/reference).  CPU only."""
import numpy as np
import torch

from conftest import load_golden
from oracle import pero_oracle as O

torch.set_num_threads(1)


def T(a):
    return torch.from_numpy(np.asarray(a))


def _replay_vq(g):
    K, decay, eps, steps = int(g["K"]), float(g["decay"]), float(g["epsilon"]), int(g["steps"])
    training = bool(int(g["training"]))
    w = T(g["weight0"])
    ema_w = T(g["ema_w0"]) if decay > 0 else None
    cs = T(g["ema_cluster_size0"]) if decay > 0 else None
    for s in range(steps):
        x = T(g[f"x{s}"])
        out = O.vq_forward(x, w, ema_w, cs, decay, eps, training)
        assert torch.equal(out["indices"], T(g[f"idx{s}"])), f"step {s}: indices"
        assert torch.equal(out["quantized"], T(g[f"q{s}"])), f"step {s}: quantized (bit-exact fp32)"
        loss = O.vq_calculate_loss(out["quantized"], x, float(g["commitment_cost"]), decay)
        np.testing.assert_allclose(float(loss), float(g[f"loss{s}"]), rtol=1e-6)
        # d(loss + <q, gq>)/dx: straight-through passes gq AND the q_latent term's gradient on tokens
        # (tokens = x + (q - x).detach() is the identity in x); commitment term acts on x as `features`.
        g_tok, g_feat = O.vq_calculate_loss_grads(out["quantized"], x, float(g["commitment_cost"]), decay)
        gx = O.vq_forward_grad_inputs(T(g[f"gq{s}"]) + g_tok) + g_feat
        np.testing.assert_allclose(gx.numpy(), g[f"gx{s}"], rtol=1e-5, atol=1e-7)
        if decay > 0 and training:
            np.testing.assert_allclose(out["ema_cluster_size"].numpy(), g[f"ema_cluster_size{s + 1}"], rtol=1e-6, atol=0)
            np.testing.assert_allclose(out["ema_w"].numpy(), g[f"ema_w{s + 1}"], rtol=1e-6, atol=1e-7)
            np.testing.assert_allclose(out["weight"].numpy(), g[f"weight{s + 1}"], rtol=2e-6, atol=1e-7)
            w, ema_w, cs = out["weight"], out["ema_w"], out["ema_cluster_size"]
        else:
            assert np.array_equal(w.numpy(), g[f"weight{s + 1}"])


def test_vq_cold_start_three_steps():
    _replay_vq(load_golden("vq_cold_3steps"))


def test_vq_warm_three_steps():
    _replay_vq(load_golden("vq_warm_3steps"))


def test_vq_no_decay():
    _replay_vq(load_golden("vq_nodecay"))


def test_vq_eval_mode_has_no_ema_side_effect():
    _replay_vq(load_golden("vq_eval"))


def test_cold_start_collapse_is_reproduced():
    """the reference's first EMA step divides by ~1e-5 for unused codes and the codebook
    collapses; the oracle must reproduce it, not 'fix' it."""
    g = load_golden("vq_cold_3steps")
    assert np.abs(g["weight1"]).max() > 1e3
    assert len(np.unique(g["idx2"])) < len(np.unique(g["idx0"]))


def test_calculate_loss_both_regimes():
    g = load_golden("vq_calculate_loss")
    for tag in ("ema", "nodecay"):
        tokens, feats, decay = T(g[f"{tag}_tokens"]), T(g[f"{tag}_features"]), float(g[f"{tag}_decay"])
        loss = O.vq_calculate_loss(tokens, feats, 0.25, decay)
        np.testing.assert_allclose(float(loss), float(g[f"{tag}_loss"]), rtol=1e-6)
        gt, gf = O.vq_calculate_loss_grads(tokens, feats, 0.25, decay, grad_out=float(g["grad_out"]))
        np.testing.assert_allclose(gt.numpy(), g[f"{tag}_g_tokens"], rtol=1e-5, atol=1e-8)
        np.testing.assert_allclose(gf.numpy(), g[f"{tag}_g_features"], rtol=1e-5, atol=1e-8)


def test_kmeans_assign_matches_reference_and_fp64():
    g = load_golden("kmeans_assign")
    feats, centers = T(g["features"]), T(g["centers"])
    f = feats.squeeze(2).permute(0, 2, 1)
    fl = f.reshape(-1, f.shape[-1])
    lab = O.kmeans_assign(fl, centers).reshape(f.shape[0], f.shape[1])
    assert np.array_equal(lab.numpy(), g["labels"])
    idx64, _, gap = O.assign_fp64(fl.numpy(), centers.numpy())
    differs = idx64 != g["labels"].reshape(-1)
    assert (gap[differs] < 1e-5).all()          # fp32 cdist vs fp64 truth may differ only at near-ties


def test_vq_indices_agree_with_fp64_outside_near_ties():
    g = load_golden("vq_warm_3steps")
    flat, _ = O.flatten_frames(T(g["x0"]))
    idx64, _, gap = O.assign_fp64(flat.numpy(), g["weight0"])
    differs = idx64 != g["idx0"]
    assert (gap[differs] < 1e-5).all()


def test_masked_ce_forward_and_gradients():
    g = load_golden("masked_ce")
    h, W, b = T(g["h"]), T(g["W"]), T(g["b"])
    labels, mask = T(g["labels"]), T(g["mask"])
    for tag, uw in (("plain", None), ("unmasked", float(g["unmasked_weight"]))):
        logits = O.linear_head(h, W, b)
        np.testing.assert_allclose(logits.numpy(), g[f"{tag}_logits"], rtol=1e-5, atol=1e-6)
        loss = O.masked_ce(logits, labels, mask, uw)
        np.testing.assert_allclose(float(loss), float(g[f"{tag}_loss"]), rtol=1e-6)
        gl = O.masked_ce_grad_logits(logits, labels, mask, uw)
        np.testing.assert_allclose(gl.numpy(), g[f"{tag}_g_logits"], rtol=1e-5, atol=1e-8)
        loss2, d_h, d_W, d_b = O.head_masked_ce(h, W, b, labels, mask, uw)
        np.testing.assert_allclose(float(loss2), float(g[f"{tag}_loss"]), rtol=1e-6)
        np.testing.assert_allclose(d_h.numpy(), g[f"{tag}_g_h"], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(d_W.numpy(), g[f"{tag}_g_W"], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(d_b.numpy(), g[f"{tag}_g_b"], rtol=1e-4, atol=1e-7)


def test_masked_ce_empty_mask_is_nan_like_reference():
    logits = torch.randn(2, 3, 8)
    labels = torch.zeros(2, 3, dtype=torch.long)
    mask = torch.zeros(2, 3, dtype=torch.long)
    assert torch.isnan(O.masked_ce(logits, labels, mask))


def test_vqvae_golden_is_self_consistent():
    """The VQVAE fixture feeds the GPU parity test; here: counts == bincount(labels), labels come from the
    projected encoder features, and the EMA touched only the vq.* state."""
    g = load_golden("vqvae_forward")
    K = g["state_vq.embedding.weight"].shape[0]
    assert np.array_equal(np.bincount(g["out_labels"], minlength=K), g["out_counts"])
    assert not np.array_equal(g["state_vq.ema_cluster_size"], g["after_vq.ema_cluster_size"])
    assert np.array_equal(g["state_encoder_projection_layer.weight"], g["after_encoder_projection_layer.weight"])


def test_create_mask_and_topk():
    rng = np.random.default_rng(0)
    labels = rng.integers(0, 10, size=(3, 20))
    labels[:, -4:] = -1
    mask = O.create_mask(labels, 0.5, np.random.default_rng(1))
    assert mask[:, -4:].sum() == 0 and set(np.unique(mask)) <= {0, 1} and mask.sum() > 0
    logits = rng.standard_normal((3, 20, 10))
    errs = O.topk_errors(logits, labels, mask, ks=(1, 3, 10))
    assert errs["errors_10"] == 0 and errs["errors_1"] >= errs["errors_3"] and errs["length"] == mask.sum()


def test_minibatch_kmeans_step_matches_sklearn_golden():
    """oracle.minibatch_kmeans_step restates scikit-learn's MiniBatchKMeans step (the fitter behind
    scripts/fit_kmeans.py:20-32); pinned on centres/counts that scikit-learn itself produced (make_golden.py)."""
    g = load_golden("kmeans_minibatch")
    c, w = g["init"], np.zeros(g["init"].shape[0], dtype=np.float32)
    for i in range(3):
        _, _, c, w = O.minibatch_kmeans_step(g[f"batch{i}"], c, w)
        np.testing.assert_allclose(w, g[f"counts{i}"], rtol=0, atol=0)
        np.testing.assert_allclose(c, g[f"centers{i}"], rtol=2e-6, atol=2e-6)


def test_pixel_mask_and_predictions_match_reference():
    """TransformerEncoder.mask (models/transformers.py:53-68) and the visualizer's argmax (visualizer.py:32) against the
    outputs of the reference's own code (tests/golden/make_golden.py gen_pixel_mask)."""
    g = load_golden("pixel_mask")
    np.testing.assert_array_equal(O.mask_tile(3, (40, 8)), g["tile"])
    np.testing.assert_array_equal(O.mask_pixels(g["x"], g["mask"], g["tile"]), g["masked"])
    np.testing.assert_array_equal(O.predict_labels(g["logits"]), g["argmax"])
    assert g["argmax"][0, 0] == 7          # the exact tie resolves to the first index


def test_vqvae_quantize_eval_matches_reference():
    """oracle.vqvae_quantize (1x1 projections as per-frame affine maps around the quantizer) against the reference's
    VQVAE.quantize in eval mode (tests/golden/make_golden.py::gen_vqvae_quantize); the fixture's labels have clear
    fp64 gaps."""
    g = load_golden("vqvae_quantize_eval")
    st = {k[len("state_"):]: T(v) for k, v in g.items() if k.startswith("state_")}
    tokens, labels = O.vqvae_quantize(T(g["features"]), st["encoder_projection_layer.weight"], st["encoder_projection_layer.bias"],
                                      st["decoder_projection_layer.weight"], st["decoder_projection_layer.bias"],
                                      st["vq.embedding.weight"])
    assert np.array_equal(labels.numpy(), g["out_labels"])
    np.testing.assert_allclose(tokens.numpy(), g["out_tokens"], rtol=1e-5, atol=1e-6)
    assert float(g["gap"].min()) > 0.05
