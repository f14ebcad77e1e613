"""Generates tests/golden/*.npz by EXECUTING THE REFERENCE modules (imported from /root/reference) on
seeded inputs, in the build container.  The reference cannot travel to the GPU box, so its inputs and
outputs are committed as small fixtures; tests/test_oracle_golden.py pins oracle/pero_oracle.py against
them and the GPU tests pin the CUDA path against them.

Run:  python tests/golden/make_golden.py          (needs /root/reference; CPU only, a few seconds)
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("PERO_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
from pero_pretraining.models.autoencoders import VQVAE, VectorQuantizer  # noqa: E402
from pero_pretraining.masked_pretraining.model import LinearHead, MaskedCrossEntropyLoss  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(1)          # fixed summation order inside the CPU GEMMs


def t2n(t):
    return t.detach().cpu().numpy().copy()


def gen_vq(name, K, D, nl, H, W, decay, steps, seed, training=True, spread=False):
    """VectorQuantizer.forward + calculate_loss + backward for `steps` consecutive calls
    (models/autoencoders.py:193-241): cold start exactly as the reference initialises."""
    torch.manual_seed(seed)
    vq = VectorQuantizer(K, D, 0.25, decay)
    vq.train(training)
    rec = {"K": K, "D": D, "decay": decay, "commitment_cost": 0.25, "epsilon": vq.epsilon, "steps": steps,
           "training": int(training)}
    if spread:   # warmed state: frames near codewords so that the index distribution stays spread
        with torch.no_grad():
            if decay > 0:
                vq.ema_w.data.copy_(vq.embedding.weight.data)
                vq.ema_cluster_size.fill_(1.0)
    rec["weight0"] = t2n(vq.embedding.weight)
    if decay > 0:
        rec["ema_w0"] = t2n(vq.ema_w)
        rec["ema_cluster_size0"] = t2n(vq.ema_cluster_size)
    g = torch.Generator().manual_seed(seed + 1)
    for s in range(steps):
        if spread:
            j = torch.randint(0, K, (nl * H * W,), generator=g)
            flat = vq.embedding.weight.data[j] + 0.5 * torch.randn(nl * H * W, D, generator=g)
            x = flat.view(nl, H, W, D).permute(0, 3, 1, 2).contiguous()
        else:
            x = torch.randn(nl, D, H, W, generator=g)
        x.requires_grad_(True)
        gq = torch.randn(nl, D, H, W, generator=g)
        q, idx = vq(x)
        loss = vq.calculate_loss(q, x)
        (loss + (q * gq).sum()).backward()
        rec[f"x{s}"] = t2n(x)
        rec[f"gq{s}"] = t2n(gq)
        rec[f"q{s}"] = t2n(q)
        rec[f"idx{s}"] = t2n(idx)
        rec[f"loss{s}"] = t2n(loss)
        rec[f"gx{s}"] = t2n(x.grad)
        rec[f"weight{s + 1}"] = t2n(vq.embedding.weight)
        if decay > 0:
            rec[f"ema_w{s + 1}"] = t2n(vq.ema_w)
            rec[f"ema_cluster_size{s + 1}"] = t2n(vq.ema_cluster_size)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)


def gen_calc_loss(name, seed):
    """calculate_loss on two unrelated same-shape tensors (the VQVAE quirk, :155-159), both decay regimes."""
    torch.manual_seed(seed)
    rec = {}
    for tag, decay in (("ema", 0.99), ("nodecay", 0.0)):
        vq = VectorQuantizer(16, 8, 0.25, decay)
        tokens = torch.randn(3, 8, 2, 5, requires_grad=True)
        feats = torch.randn(3, 8, 2, 5, requires_grad=True)
        loss = vq.calculate_loss(tokens, feats)
        (loss * 1.7).backward()
        rec[f"{tag}_tokens"] = t2n(tokens)
        rec[f"{tag}_features"] = t2n(feats)
        rec[f"{tag}_loss"] = t2n(loss)
        rec[f"{tag}_g_tokens"] = t2n(tokens.grad) if tokens.grad is not None else np.zeros_like(t2n(tokens))
        rec[f"{tag}_g_features"] = t2n(feats.grad)
        rec[f"{tag}_decay"] = decay
    rec["grad_out"] = 1.7
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)


class _Enc(torch.nn.Module):
    out_channels = 6

    def __init__(self):
        super().__init__()
        self.conv = torch.nn.Conv2d(3, 6, (4, 8), stride=(4, 8))

    def forward(self, x):
        return self.conv(x)


class _Dec(torch.nn.Module):
    base_channels = 6

    def __init__(self):
        super().__init__()
        self.conv = torch.nn.ConvTranspose2d(6, 3, (4, 8), stride=(4, 8))

    def forward(self, x):
        return self.conv(x)


def gen_vqvae(name, seed):
    """VQVAE.forward end to end with stand-in conv encoder/decoder (models/autoencoders.py:148-167): pins the
    projection / calculate_loss(tokens, features) wiring and the `counts` output."""
    torch.manual_seed(seed)
    m = VQVAE(_Enc(), _Dec(), num_embeddings=32, embeddings_dim=8)
    m.train()
    rec = {"state_" + k: t2n(v) for k, v in m.state_dict().items()}
    images = torch.rand(4, 3, 4, 64)
    out = m(images)
    out["loss"].backward()
    rec["images"] = t2n(images)
    for k in ("tokens", "labels", "loss", "reconstructions", "counts"):
        rec["out_" + k] = t2n(out[k])
    rec["grad_enc_proj_w"] = t2n(m.encoder_projection_layer.weight.grad)
    rec["grad_dec_proj_w"] = t2n(m.decoder_projection_layer.weight.grad)
    rec["grad_enc_w"] = t2n(m.encoder.conv.weight.grad)
    for k, v in m.state_dict().items():
        rec["after_" + k] = t2n(v)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)


def gen_kmeans(name, seed):
    """scripts/produce_kmeans_labels.py:34, 72-80 restated verbatim on the reference's torch ops (the script
    itself needs lmdb/safe_gpu, which are not installed)."""
    g = torch.Generator().manual_seed(seed)
    B, D, T, K = 5, 24, 17, 40
    kmeans_model = torch.randn(K, D, generator=g)
    features = torch.randn(B, D, 1, T, generator=g)
    km = kmeans_model.reshape(1, kmeans_model.shape[0], kmeans_model.shape[1])
    f = features.squeeze(2)
    f = f.permute(0, 2, 1)
    fl = f.reshape(-1, f.shape[-1])
    distances = torch.cdist(fl, km).squeeze()
    assignment = torch.argmin(distances, dim=1)
    assignment = assignment.reshape(f.shape[0], f.shape[1])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), centers=t2n(kmeans_model), features=t2n(features),
                        labels=t2n(assignment), distances=t2n(distances))


def gen_masked_ce(name, seed):
    """LinearHead + MaskedCrossEntropyLoss fwd/bwd (masked_pretraining/model.py:72-105), with -1 padded labels,
    with and without unmasked_weight."""
    torch.manual_seed(seed)
    Nl, T, Dh, V = 4, 24, 32, 96
    head = LinearHead(Dh, V)
    rng = np.random.default_rng(seed)
    h = torch.randn(Nl, T, Dh, requires_grad=True)
    labels = torch.from_numpy(rng.integers(0, V, size=(Nl, T))).long()
    labels[:, -5:] = -1
    mask_np = (rng.random((Nl, T)) < 0.3).astype(int) * (labels.numpy() >= 0).astype(int)
    mask = torch.from_numpy(mask_np)
    rec = {"h": t2n(h), "W": t2n(head.linear.weight), "b": t2n(head.linear.bias), "labels": t2n(labels), "mask": mask_np}
    for tag, uw in (("plain", None), ("unmasked", 0.3)):
        for p in (h, head.linear.weight, head.linear.bias):
            p.grad = None
        logits = head(h)
        loss = MaskedCrossEntropyLoss(uw)(logits, labels, mask)
        logits.retain_grad()
        loss.backward()
        rec[f"{tag}_logits"] = t2n(logits)
        rec[f"{tag}_loss"] = t2n(loss)
        rec[f"{tag}_g_logits"] = t2n(logits.grad)
        rec[f"{tag}_g_h"] = t2n(h.grad)
        rec[f"{tag}_g_W"] = t2n(head.linear.weight.grad)
        rec[f"{tag}_g_b"] = t2n(head.linear.bias.grad)
    rec["unmasked_weight"] = 0.3
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)


def gen_kmeans_minibatch(name, seed):
    """Third-party arithmetic of the codebook fit (scripts/fit_kmeans.py:20-32): scikit-learn's own
    MiniBatchKMeans.partial_fit, from given centres, on three given batches, reassignment off
    (reassignment_ratio=0 — the random part is not parity material).  Well separated blobs: no near-tie labels."""
    import sklearn
    from sklearn.cluster import MiniBatchKMeans
    rng = np.random.RandomState(seed)
    K, D, nb = 24, 20, 160
    true = rng.randn(K, D).astype(np.float32) * 4.0
    init = (true + 0.3 * rng.randn(K, D)).astype(np.float32)
    batches = [(true[rng.randint(0, K, nb)] + 0.4 * rng.randn(nb, D)).astype(np.float32) for _ in range(3)]
    km = MiniBatchKMeans(n_clusters=K, init=init, n_init=1, batch_size=nb, reassignment_ratio=0.0, compute_labels=True,
                         random_state=0)
    rec = {"init": init, "sklearn_version": np.array(sklearn.__version__)}
    for i, b in enumerate(batches):
        km.partial_fit(b)
        rec[f"batch{i}"] = b
        rec[f"centers{i}"] = km.cluster_centers_.astype(np.float32).copy()
        rec[f"counts{i}"] = km._counts.astype(np.float32).copy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)


def gen_pixel_mask(name, seed):
    """TransformerEncoder.mask (models/transformers.py:53-68) executed from the reference's own code.  The class cannot
    be constructed on a CPU-only host (its __init__ moves the pattern .to("cuda"), :34), so the unbound method runs on
    a stand-in object that carries exactly the attributes the method reads, built the way __init__ builds them
    (:27-34) but left on the CPU.  Also stores argmax labels of a small logits tensor (visualizer.py:32)."""
    import types
    from pero_pretraining.models.transformers import TransformerEncoder
    rng = np.random.RandomState(seed)
    N, C, H, W, pw = 3, 3, 40, 136, 8
    np.random.seed(42)                                                                     # :30
    mask_tile = torch.tensor(np.random.rand(1, C, 40, pw), dtype=torch.float32)              # :31-32
    stand_in = types.SimpleNamespace(in_channels=C, height=H, mask_pattern=mask_tile.repeat(1, 1, 1, 512))   # :34 minus .to("cuda")
    x = torch.tensor(rng.rand(N, C, H, W), dtype=torch.float32)
    mask = (rng.rand(N, W // pw) < 0.3).astype(np.int64)
    mask[0, 0] = 1
    mask[2, -1] = 1
    out = TransformerEncoder.mask(stand_in, x.clone(), mask)
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(2, 9, 50, generator=g)
    logits[0, 0, 7] = logits[0, 0, 31] = 9.0                                                 # exact tie: first index wins
    rec = {"x": t2n(x), "mask": mask, "masked": t2n(out), "tile": t2n(mask_tile[0]),
           "logits": t2n(logits), "argmax": t2n(torch.argmax(logits, dim=-1))}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)


class _Pass(torch.nn.Module):
    def __init__(self, c):
        super().__init__()
        self.out_channels = self.base_channels = c

    def forward(self, x):
        return x


def gen_vqvae_quantize(name, seed, C=40, D=24, K=96, nl=3, H=2, W=21):
    """VQVAE.quantize in eval mode (label production, scripts/produce_vqvae_labels.py:37): the reference's two 1x1
    convolutions around its quantizer on features whose projection lies near a codeword (so that the labels have clear
    fp64 gaps); also the fp64 top-2 gap of every frame for the near-tie rule."""
    torch.manual_seed(seed)
    m = VQVAE(_Pass(C), _Pass(C), num_embeddings=K, embeddings_dim=D)
    m.eval()
    with torch.no_grad():
        we = m.encoder_projection_layer.weight.view(D, C).double()
        N = nl * H * W
        target = m.vq.embedding.weight[torch.randint(0, K, (N,))].double() + 0.3 * torch.randn(N, D).double()
        rows = (target - m.encoder_projection_layer.bias.double()) @ torch.linalg.pinv(we).t()
        feats = rows.float().view(nl, H, W, C).permute(0, 3, 1, 2).contiguous()
        tokens, labels = m.quantize(feats)
        xp = m.encoder_projection_layer(feats).permute(0, 2, 3, 1).reshape(-1, D).double()
        d = torch.cdist(xp, m.vq.embedding.weight.double()) ** 2
        top2 = torch.topk(d, 2, dim=1, largest=False).values
        gap = (top2[:, 1] - top2[:, 0]) / top2[:, 1].clamp_min(1e-30)
    rec = {"state_" + k: t2n(v) for k, v in m.state_dict().items()}
    rec.update(features=t2n(feats), out_tokens=t2n(tokens), out_labels=t2n(labels), gap=t2n(gap),
               dims=np.array([C, D, K, nl, H, W]))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "vqvae_quantize":
        gen_vqvae_quantize("vqvae_quantize_eval", seed=81)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "kmeans_minibatch":       # only the scikit-learn fixture
        gen_kmeans_minibatch("kmeans_minibatch", seed=61)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "pixel_mask":
        gen_pixel_mask("pixel_mask", seed=71)
        sys.exit(0)
    gen_vq("vq_cold_3steps", K=64, D=16, nl=4, H=1, W=12, decay=0.99, steps=3, seed=11)
    gen_vq("vq_warm_3steps", K=48, D=32, nl=3, H=2, W=20, decay=0.99, steps=3, seed=12, spread=True)
    gen_vq("vq_nodecay", K=32, D=8, nl=2, H=1, W=16, decay=0.0, steps=1, seed=13)
    gen_vq("vq_eval", K=64, D=16, nl=4, H=1, W=12, decay=0.99, steps=1, seed=14, training=False)
    gen_calc_loss("vq_calculate_loss", seed=21)
    gen_vqvae("vqvae_forward", seed=31)
    gen_kmeans("kmeans_assign", seed=41)
    gen_masked_ce("masked_ce", seed=51)
    gen_kmeans_minibatch("kmeans_minibatch", seed=61)
    gen_pixel_mask("pixel_mask", seed=71)
    gen_vqvae_quantize("vqvae_quantize_eval", seed=81)
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f"  {f}: {os.path.getsize(os.path.join(OUT, f))} bytes")
