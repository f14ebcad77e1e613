"""CPU oracle for the quantize-and-predict path of DCGM/pero-pretraining.

THIS IS TEST INFRASTRUCTURE.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it; nothing under ``pero_pretraining_b200/`` does,
and the product path fails loudly when the CUDA library is missing instead of coming here.

What it is: a functional restatement, in my own words, of the reference's arithmetic for this path.
The reference's "engine" for the path is PyTorch itself (every op on the path is a
torch library call; there is no third-party kernel to restate), so the restatement keeps the
reference's op ORDER in fp32 torch-CPU ops -- fp32 results depend on that order -- and adds
 (i) explicit backward formulas (no autograd) so gradients are pinned independently, and
 (ii) an fp64 brute-force assignment that yields the ground-truth index and the top-2 gap used by the
      near-tie rule of the parity tests.

Parity pinning: the reference ships NO tests, golden vectors or fixtures for this path
("parity unpinned by the reference's own tests").  The oracle is therefore pinned against
outputs of the reference modules themselves, executed in the build container by
``tests/golden/make_golden.py`` (which imports /root/reference) and committed as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` replays them bit-for-bit / to 1e-6.

Citations are file:line relative to the reference root.
"""
import numpy as np
import torch


# ------------------------------------------------------------------------------------------ VQ forward
def flatten_frames(inputs):
    """[Nl, D, H, W] -> ([N, D] rows, NHWC shape).  models/autoencoders.py:205-209."""
    nhwc = inputs.permute(0, 2, 3, 1).contiguous()
    return nhwc.view(-1, nhwc.shape[-1]), nhwc.shape


def vq_distances(flat, weight):
    """|x|^2 + |c|^2 - 2 x.c^T in the reference's association order.  models/autoencoders.py:212-214."""
    return (flat ** 2).sum(dim=1, keepdim=True) + (weight ** 2).sum(dim=1) - 2 * torch.matmul(flat, weight.t())


def vq_assign_fp32(flat, weight):
    """argmin over codewords, first index on ties.  models/autoencoders.py:217."""
    return torch.argmin(vq_distances(flat, weight), dim=1)


def vq_forward(inputs, weight, ema_w=None, ema_cluster_size=None, decay=0.99, epsilon=1e-5, training=True,
               indices_override=None):
    """VectorQuantizer.forward restated as a pure function.  models/autoencoders.py:204-241.

    Returns dict(quantized [Nl,D,H,W] (forward value of the straight-through expression), indices [N] int64,
    and -- when decay > 0 and training -- the NEW weight / ema_w / ema_cluster_size; the quantized output of
    this call uses the OLD weight, exactly as the reference (the update takes effect next step).
    `indices_override` replaces the arg-min result (tests use it to check everything downstream of the
    assignment independently of bf16 near-tie flips)."""
    flat, nhwc_shape = flatten_frames(inputs)
    K = weight.shape[0]
    idx = vq_assign_fp32(flat, weight) if indices_override is None else indices_override
    # :218-222 one-hot @ weight  ==  gather (each output row has exactly one non-zero product)
    encodings = torch.zeros(idx.shape[0], K, dtype=flat.dtype)
    encodings.scatter_(1, idx.unsqueeze(1), 1)
    quantized = torch.matmul(encodings, weight).view(nhwc_shape)
    out = {"indices": idx}
    if decay > 0.0 and training:
        counts = encodings.sum(0)                                          # :226
        cs = ema_cluster_size * decay + (1 - decay) * counts              # :226
        n = cs.sum()                                                       # :229
        cs = (cs + epsilon) / (n + K * epsilon) * n                        # :230-232
        dw = torch.matmul(encodings.t(), flat)                             # :234
        new_ema_w = ema_w * decay + (1 - decay) * dw                       # :235
        out.update(weight=new_ema_w / cs.unsqueeze(1), ema_w=new_ema_w, ema_cluster_size=cs,   # :237
                   counts=counts, dw=dw)
    nhwc = inputs.permute(0, 2, 3, 1).contiguous()
    st = nhwc + (quantized - nhwc)                                         # :239 forward value
    out["quantized"] = st.permute(0, 3, 1, 2).contiguous()                 # :241
    return out


def vq_forward_grad_inputs(grad_quantized):
    """Straight-through: d quantized / d inputs = identity, nothing flows to the codebook (:239)."""
    return grad_quantized.clone()


def vq_calculate_loss(tokens, features, commitment_cost, decay):
    """VectorQuantizer.calculate_loss.  models/autoencoders.py:193-202."""
    e_latent = ((tokens - features) ** 2).mean()
    q_latent = e_latent if not decay > 0.0 else 0.0
    return q_latent + commitment_cost * e_latent


def vq_calculate_loss_grads(tokens, features, commitment_cost, decay, grad_out=1.0):
    """(d loss / d tokens, d loss / d features).  mse(tokens.detach(), features) only reaches `features`;
    the q_latent term (decay == 0) only reaches `tokens` (:195-200)."""
    diff = (features - tokens) * (2.0 / tokens.numel()) * grad_out
    g_features = commitment_cost * diff
    g_tokens = -diff if not decay > 0.0 else torch.zeros_like(tokens)
    return g_tokens, g_features


def conv1x1(x, weight, bias):
    """torch.nn.Conv2d(C_in, C_out, 1) as the per-frame affine map it is: x [Nl, C, H, W], weight [C_out, C_in(, 1, 1)],
    bias [C_out] -> [Nl, C_out, H, W].  models/autoencoders.py:114-115."""
    w = weight.reshape(weight.shape[0], -1)
    y = torch.matmul(x.permute(0, 2, 3, 1), w.t())
    if bias is not None:
        y = y + bias
    return y.permute(0, 3, 1, 2).contiguous()


def vqvae_quantize(features, enc_w, enc_b, dec_w, dec_b, weight, indices_override=None):
    """VQVAE.quantize in eval mode, models/autoencoders.py:142-147: encoder projection -> VectorQuantizer.forward (no EMA
    update) -> decoder projection.  Returns (projected tokens [Nl, Cd, H, W], labels [N])."""
    x = conv1x1(features, enc_w, enc_b)
    out = vq_forward(x, weight, training=False, indices_override=indices_override)
    return conv1x1(out["quantized"], dec_w, dec_b), out["indices"]


def bincount(labels, K):
    """VQVAE.forward 'counts'.  models/autoencoders.py:165."""
    return torch.bincount(labels, minlength=K)


# ------------------------------------------------------------------------------------------ FQ / PQ-AE
def kmeans_assign(features_linear, centers):
    """Nearest-centre labels of the k-means labeller: cdist (Euclidean) + argmin.
    scripts/produce_kmeans_labels.py:34, 72-76."""
    d = torch.cdist(features_linear, centers.reshape(1, centers.shape[0], centers.shape[1])).squeeze()
    if d.dim() == 1:      # a single frame: squeeze() dropped the row axis as well
        d = d.reshape(features_linear.shape[0], -1)
    return torch.argmin(d, dim=1)


# ------------------------------------------------------------------------------------------ fp64 truth
def assign_fp64(flat, weight, chunk=4096):
    """Brute-force fp64 assignment: (idx, d_min, relative top-2 gap).  Ground truth for the near-tie
    rule: a CUDA index may differ from the reference only where `gap` < the epsilon stated in the test."""
    x = np.asarray(flat, dtype=np.float64)
    c = np.asarray(weight, dtype=np.float64)
    cn = (c * c).sum(1)
    idx = np.empty(x.shape[0], dtype=np.int64)
    dmin = np.empty(x.shape[0])
    gap = np.empty(x.shape[0])
    for s in range(0, x.shape[0], chunk):
        xs = x[s:s + chunk]
        d = (xs * xs).sum(1, keepdims=True) + cn[None, :] - 2.0 * xs @ c.T
        order = np.argsort(d, axis=1, kind="stable")[:, :2]
        rows = np.arange(xs.shape[0])
        d0 = d[rows, order[:, 0]]
        d1 = d[rows, order[:, 1]] if c.shape[0] > 1 else np.full_like(d0, np.inf)
        idx[s:s + chunk] = order[:, 0]
        dmin[s:s + chunk] = d0
        gap[s:s + chunk] = (d1 - d0) / np.maximum(np.abs(d0), 1e-30)
    return idx, dmin, gap


def assign_fp64_torch(flat, weight, chunk=2048):
    """The same fp64 brute force as assign_fp64, stated in torch ops so that the TEST may run it on whatever
    device the tensors live on (a B200 does the 8192 x 8192 x 256 case in milliseconds, which is what lets the
    near-tie rule be applied at the BASELINE sizes; models/autoencoders.py:212-217 in fp64).  Returns torch tensors
    (idx int64, d_min fp64, relative top-2 gap fp64) on flat's device.  Ties: lowest index, like torch.argmin."""
    x = flat.double()
    c = weight.double()
    cn = (c * c).sum(1)
    N, K = x.shape[0], c.shape[0]
    idx = torch.empty(N, dtype=torch.int64, device=x.device)
    dmin = torch.empty(N, dtype=torch.float64, device=x.device)
    gap = torch.empty(N, dtype=torch.float64, device=x.device)
    for s in range(0, N, chunk):
        xs = x[s:s + chunk]
        d = (xs * xs).sum(1, keepdim=True) + cn[None, :] - 2.0 * (xs @ c.t())
        i0 = torch.argmin(d, dim=1)                             # first minimal index
        d0 = d.gather(1, i0[:, None]).squeeze(1)
        if K > 1:
            d.scatter_(1, i0[:, None], float("inf"))
            d1 = d.min(dim=1).values
        else:
            d1 = torch.full_like(d0, float("inf"))
        idx[s:s + chunk] = i0
        dmin[s:s + chunk] = d0
        gap[s:s + chunk] = (d1 - d0) / d0.abs().clamp_min(1e-30)
    return idx, dmin, gap


# ------------------------------------------------------------------------------------------ masked CE
def linear_head(x, W, b):
    """LinearHead.forward.  masked_pretraining/model.py:104-105."""
    return torch.nn.functional.linear(x, W, b)


def _ce_mean(logits, labels):
    """mean_i (logsumexp(z_i) - z_i[label_i]); NaN on an empty selection like F.cross_entropy."""
    lse = torch.logsumexp(logits.float(), dim=1)
    picked = logits.float().gather(1, labels.unsqueeze(1)).squeeze(1)
    return (lse - picked).mean()


def masked_ce(output, labels, mask, unmasked_weight=None):
    """MaskedCrossEntropyLoss.forward.  masked_pretraining/model.py:78-95."""
    sel = mask == 1
    loss = _ce_mean(output[sel], labels[sel])
    if unmasked_weight is not None:
        un = mask == 0                                           # :85-86
        un_out, un_lab = output[un], labels[un]
        keep = un_lab >= 0                                       # :88-90 drops the -1 padding
        loss = loss + unmasked_weight * _ce_mean(un_out[keep], un_lab[keep])
    return loss


def masked_ce_grad_logits(output, labels, mask, unmasked_weight=None, grad_out=1.0):
    """d loss / d output: (softmax - onehot) / M on the selected frames, zero elsewhere."""
    g = torch.zeros_like(output, dtype=torch.float32)

    def add(sel, scale):
        m = int(sel.sum())
        if m == 0:
            return
        z = output[sel].float()
        p = torch.softmax(z, dim=1)
        p[torch.arange(m), labels[sel]] -= 1.0
        g[sel] += p * (scale / m)

    add(mask == 1, grad_out)
    if unmasked_weight is not None:
        add((mask == 0) & (labels >= 0), grad_out * unmasked_weight)
    return g


def head_masked_ce(h, W, b, labels, mask, unmasked_weight=None):
    """LinearHead + MaskedCrossEntropyLoss with explicit gradients: the fused CUDA path's contract.
    Returns loss, d_h [like h], d_W, d_b.  masked_pretraining/model.py:49, 60-61, 78-95, 104-105."""
    logits = linear_head(h.float(), W, b)
    loss = masked_ce(logits, labels, mask, unmasked_weight)
    g = masked_ce_grad_logits(logits, labels, mask, unmasked_weight)
    g2 = g.reshape(-1, g.shape[-1])
    d_h = (g2 @ W).reshape(h.shape)
    d_W = g2.t() @ h.float().reshape(-1, h.shape[-1])
    d_b = g2.sum(0)
    return loss, d_h, d_W, d_b


def create_mask(labels, masking_prob, rng):
    """BatchOperator._create_mask: host-side numpy mask, zero on -1 padding.
    masked_pretraining/batch_operator.py:27-32 (the reference draws from the global numpy RNG;
    here the generator is explicit so that tests are seeded)."""
    active = (labels >= 0).astype(int)
    return (rng.random(labels.shape) < masking_prob).astype(int) * active


def mask_pixels(x, mask, tile):
    """TransformerEncoder.mask: the 8-px image column of every masked frame is overwritten with the fixed noise tile.
    models/transformers.py:53-68 (pattern = tile repeated along the width, :34).  x [N, C, H, W], mask [N, W/8] {0,1},
    tile [C, H, pw]; returns a new array."""
    x = np.array(x, dtype=np.float32, copy=True)
    mask = np.asarray(mask)
    pw = tile.shape[2]
    for n, t in zip(*np.nonzero(mask == 1)):
        w0, w1 = t * pw, min((t + 1) * pw, x.shape[3])
        x[n, :, :, w0:w1] = tile[:, :, :w1 - w0]
    return x


def mask_tile(in_channels=3, patch_size=(40, 8)):
    """The reference's fixed noise tile: np.random.seed(42); np.random.rand(1, C, ph, pw).  models/transformers.py:29-32."""
    return np.random.RandomState(42).rand(1, in_channels, patch_size[0], patch_size[1]).astype(np.float32)[0]


def predict_labels(logits):
    """MaskedVisualizer predictions: argmax over the label axis, first index on ties.  masked_pretraining/visualizer.py:32."""
    return np.argmax(np.asarray(logits), axis=-1)


def topk_errors(logits, labels, mask, ks=(1, 3, 10)):
    """Tester._update_errors: top-k error counts on masked frames.  masked_pretraining/tester.py:70-93."""
    sel = mask == 1
    z = np.asarray(logits)[sel]
    y = np.asarray(labels)[sel]
    out = {"length": int(sel.sum())}
    order = np.argsort(-z, axis=1, kind="stable")
    for k in ks:
        hit = (order[:, :k] == y[:, None]).any(1)
        out[f"errors_{k}"] = int((~hit).sum())
    return out


def minibatch_kmeans_step(X, centers, weight_sums):
    """One step of scikit-learn's MiniBatchKMeans without reassignment (the fitter of scripts/fit_kmeans.py:20-32;
    scikit-learn is an unpinned dependency of the reference — 1.9.0 in the build container).  Restates
    sklearn/cluster/_kmeans.py _mini_batch_step + _k_means_minibatch.pyx update_center_dense in fp32:
        labels = nearest centre (squared Euclidean);  inertia = sum of squared distances BEFORE the update
        for centres with members: c <- (c * w + sum_{members, in sample order} x) * (1 / (w + n));  w <- w + n
    Returns (labels int64 [n], inertia float, centers_new [K, D] fp32, weight_sums_new [K] fp32)."""
    X = np.asarray(X, dtype=np.float32)
    c = np.asarray(centers, dtype=np.float32).copy()
    w = np.asarray(weight_sums, dtype=np.float32).copy()
    d = (X.astype(np.float64) ** 2).sum(1)[:, None] - 2.0 * X.astype(np.float64) @ c.astype(np.float64).T + (c.astype(np.float64) ** 2).sum(1)[None]
    labels = d.argmin(1)
    inertia = float(d[np.arange(len(X)), labels].sum())
    for k in np.unique(labels):
        members = X[labels == k]
        acc = c[k] * w[k]
        for row in members:                       # sample order, fp32 adds: what the Cython loop does
            acc = acc + row
        w[k] = w[k] + np.float32(len(members))
        c[k] = acc * (np.float32(1.0) / w[k])
    return labels.astype(np.int64), inertia, c, w
