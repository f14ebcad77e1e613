"""Tensor-level wrappers over the C ABI (include/pero_b200.h).

Each function takes/returns torch CUDA tensors, passes raw device pointers + sizes + the current CUDA
stream to libpero_b200.so and raises on any non-zero return code.  PyTorch is used for device memory and
streams only; no torch op computes anything on the path here.
"""
import torch

from . import _lib
from ._lib import check


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    """Raw handle of the current CUDA stream of the current device (the cheap accessor when torch has it:
    torch.cuda.current_stream() costs several microseconds per call, and every op here needs the handle)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _f32c(t, name):
    if not (t.is_cuda and t.dtype == torch.float32):
        raise TypeError(f"{name} must be a CUDA float32 tensor, got {t.dtype} on {t.device}")
    return t if t.is_contiguous() else t.contiguous()


def _ws(nbytes, device):
    # 256-byte aligned scratch from torch's caching allocator (its blocks are 512-byte aligned)
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def require_device():
    """Fail loudly unless a B200 (sm_100) is current and the library is loadable."""
    L = _lib.lib()
    if not torch.cuda.is_available():
        raise _lib.PeroError("pero_pretraining_b200 needs a CUDA device (B200); there is no CPU fallback")
    check(L.pero_check_device(), "pero_check_device")


# ------------------------------------------------------------------------------------------ codebook / assign
class PreparedCodebook:
    """Device blob with the bf16 GEMM operand and |c|^2 of a [K, D] fp32 codebook."""

    def __init__(self, K, D, device):
        self.K, self.D = int(K), int(D)
        self.nbytes = _lib.lib().pero_vq_codebook_bytes(self.K, self.D)
        self.blob = _ws(self.nbytes, device)

    def prepare(self, weight):
        w = _f32c(weight, "weight")
        assert tuple(w.shape) == (self.K, self.D)
        check(_lib.lib().pero_vq_codebook_prepare(w.data_ptr(), self.K, self.D, self.blob.data_ptr(), self.nbytes,
                                                  _stream()), "pero_vq_codebook_prepare")
        return self


def vq_assign(x, codebook, n_lines, frames_per_line, channels_first, want_dmin=False, want_rows=False,
              index_offset=0, packed=None, init_packed=False):
    """Nearest-codeword indices.  x: [n_lines, D, frames] (channels_first) or [N, D].
    Returns (idx int64 [N] or None, dmin or None, x_rows or None).
    packed: min-merge the packed (distance, index) winners into this int64 [N] buffer instead (pre-set by vq_packed_init,
    or reset by this call's frame preparation pass when init_packed)."""
    L = _lib.lib()
    x = _f32c(x, "x")
    N = int(n_lines) * int(frames_per_line)
    K, D = codebook.K, codebook.D
    dev = x.device
    idx = torch.empty(N, dtype=torch.int64, device=dev) if packed is None else None
    dmin = torch.empty(N, dtype=torch.float32, device=dev) if (want_dmin and packed is None) else None
    x_rows = torch.empty(N, D, dtype=torch.float32, device=dev) if want_rows else None
    if N == 0:
        return idx, dmin, x_rows
    wsb = L.pero_vq_assign_workspace_bytes(N, K, D)
    ws = _ws(wsb, dev)
    check(L.pero_vq_assign(x.data_ptr(), int(n_lines), int(frames_per_line), (1 if channels_first else 0) | (2 if init_packed else 0), K, D,
                           codebook.blob.data_ptr(), int(index_offset), _p(idx), _p(dmin), _p(packed), _p(x_rows),
                           ws.data_ptr(), wsb, _stream()), "pero_vq_assign")
    return idx, dmin, x_rows


def vq_forward(x, codebook, weight, ema_w, ema_cluster_size, decay, epsilon, update_ema, n_lines, frames_per_line,
               channels_first=True, out=None, idx=None, ws=None):
    """The whole VectorQuantizer.forward in one call of the C ABI (pero_vq_forward): returns (quantized with the
    layout of x, idx int64 [N]); with update_ema the EMA state, weight and the prepared codebook are updated in place.
    out / idx / ws: preallocated outputs and workspace (CUDA-graph capture with static buffers)."""
    L = _lib.lib()
    N = int(n_lines) * int(frames_per_line)
    K, D = codebook.K, codebook.D
    dev = x.device
    out = torch.empty_like(x) if out is None else out
    idx = torch.empty(N, dtype=torch.int64, device=dev) if idx is None else idx
    if N == 0:
        return out, idx
    upd = 1 if update_ema else 0
    wsb = L.pero_vq_forward_workspace_bytes(N, K, D, upd)
    if ws is None or ws.numel() < wsb:
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    check(L.pero_vq_forward(x.data_ptr(), int(n_lines), int(frames_per_line), 1 if channels_first else 0, K, D,
                            codebook.blob.data_ptr(), codebook.nbytes, weight.data_ptr(), _p(ema_w), _p(ema_cluster_size),
                            float(decay), float(epsilon), upd, out.data_ptr(), idx.data_ptr(), ws.data_ptr(), wsb, _stream()),
          "pero_vq_forward")
    return out, idx


def vq_forward_workspace(N, K, D, update_ema, device):
    return torch.empty(_lib.lib().pero_vq_forward_workspace_bytes(int(N), int(K), int(D), 1 if update_ema else 0), dtype=torch.uint8,
                       device=device)


def vq_packed_init(N, device, out=None):
    packed = out if out is not None else torch.empty(int(N), dtype=torch.int64, device=device)
    check(_lib.lib().pero_vq_packed_init(packed.data_ptr(), int(N), _stream()), "pero_vq_packed_init")
    return packed


def vq_unpack(packed, want_dmin=False):
    N = packed.numel()
    idx = torch.empty(N, dtype=torch.int64, device=packed.device)
    dmin = torch.empty(N, dtype=torch.float32, device=packed.device) if want_dmin else None
    check(_lib.lib().pero_vq_unpack(packed.data_ptr(), N, idx.data_ptr(), _p(dmin), _stream()), "pero_vq_unpack")
    return idx, dmin


def vq_assign_bf16(x_bf16, codebook, packed, index_offset=0):
    """The distance GEMM + arg-min on frames that already exist as bf16 rows [N, Dp] (pero_vq_assign_bf16): min-merges
    the packed (distance, index) winners into `packed` (pre-set to "empty")."""
    N = x_bf16.shape[0]
    check(_lib.lib().pero_vq_assign_bf16(x_bf16.data_ptr(), int(N), codebook.K, codebook.D, codebook.blob.data_ptr(),
                                         int(index_offset), packed.data_ptr(), _stream()), "pero_vq_assign_bf16")
    return packed


def proj_forward(x, weight, bias, n_lines, frames_per_line, channels_first, want_rows=True, want_bf16=False, packed=None):
    """1x1 projection y[n, :] = weight x[n, :] + bias over the N = n_lines * frames_per_line frames (pero_proj_forward).
    x: [n_lines, C, frames] fp32 (channels_first) or rows [N, C]; weight [D, C]; bias [D] or None.
    Returns (rows fp32 [N, D] or None, bf16 rows [N, Dp] or None: the distance GEMM's operand); `packed` (int64 [N]) is
    reset to "empty" on the way."""
    L = _lib.lib()
    x = _f32c(x, "x")
    w = _f32c(weight, "weight")
    b = None if bias is None else _f32c(bias, "bias")
    D, C = int(w.shape[0]), int(w.shape[1])
    N = int(n_lines) * int(frames_per_line)
    dev = x.device
    rows = torch.empty(N, D, dtype=torch.float32, device=dev) if want_rows else None
    xb = torch.empty(N, (D + 63) // 64 * 64, dtype=torch.bfloat16, device=dev) if want_bf16 else None
    if N == 0:
        return rows, xb
    wsb = L.pero_proj_workspace_bytes(N, C, D)
    ws = _ws(wsb, dev)
    check(L.pero_proj_forward(x.data_ptr(), int(n_lines), int(frames_per_line), 1 if channels_first else 0, C, w.data_ptr(), _p(b), D,
                              _p(rows), _p(xb), _p(packed), ws.data_ptr(), wsb, _stream()), "pero_proj_forward")
    return rows, xb


def gather_rows_cf(table, idx, n_lines, frames_per_line):
    """out[l, c, t] = table[idx[l * frames + t], c] as a channels-first tensor [n_lines, C, frames] (pero_gather_rows_cf)."""
    t = _f32c(table, "table")
    K, C = int(t.shape[0]), int(t.shape[1])
    out = torch.empty(int(n_lines), C, int(frames_per_line), dtype=torch.float32, device=t.device)
    check(_lib.lib().pero_gather_rows_cf(t.data_ptr(), idx.data_ptr(), int(n_lines), int(frames_per_line), K, C, out.data_ptr(),
                                         _stream()), "pero_gather_rows_cf")
    return out


def vq_gather_st(x_rows, idx, weight, n_lines, frames_per_line, channels_first):
    """out = x + (weight[idx] - x), channels-first [n_lines, D, frames] or rows [N, D]."""
    w = _f32c(weight, "weight")
    K, D = w.shape
    shape = (int(n_lines), D, int(frames_per_line)) if channels_first else (int(n_lines) * int(frames_per_line), D)
    out = torch.empty(shape, dtype=torch.float32, device=w.device)
    check(_lib.lib().pero_vq_gather_st(x_rows.data_ptr(), idx.data_ptr(), w.data_ptr(), int(n_lines),
                                       int(frames_per_line), 1 if channels_first else 0, K, D, out.data_ptr(),
                                       _stream()), "pero_vq_gather_st")
    return out


def vq_gather_st_mse(x_rows, idx, weight, n_lines, frames_per_line, channels_first, scale_a=1.0, scale_b=0.0):
    """(out = x + (weight[idx] - x), the 0-dim loss scale_a * m + scale_b * m with m = mean((out - x)^2)) in one pass
    (pero_vq_gather_st_mse): VectorQuantizer.forward's output and calculate_loss's value without re-reading either."""
    L = _lib.lib()
    w = _f32c(weight, "weight")
    K, D = w.shape
    cf = 1 if channels_first else 0
    shape = (int(n_lines), D, int(frames_per_line)) if channels_first else (int(n_lines) * int(frames_per_line), D)
    out = torch.empty(shape, dtype=torch.float32, device=w.device)
    loss = torch.empty((), dtype=torch.float32, device=w.device)
    wsb = L.pero_vq_gather_st_mse_workspace_bytes(int(n_lines), int(frames_per_line), cf, D)
    ws = _ws(wsb, w.device)
    check(L.pero_vq_gather_st_mse(x_rows.data_ptr(), idx.data_ptr(), w.data_ptr(), int(n_lines), int(frames_per_line), cf, K, D,
                                  out.data_ptr(), float(scale_a), float(scale_b), loss.data_ptr(), ws.data_ptr(), wsb, _stream()),
          "pero_vq_gather_st_mse")
    return out, loss


def vq_ema_accumulate(x_rows, idx, K, out=None):
    """Deterministic per-codeword sums and counts: one fp32 buffer [K*D + K] (`out`: e.g. a peer-buffer range)."""
    L = _lib.lib()
    N, D = x_rows.shape
    if out is None:
        out = torch.empty(K * D + K, dtype=torch.float32, device=x_rows.device)
    elif out.numel() != K * D + K or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("out must be a contiguous float32 buffer of K*D + K elements")
    wsb = L.pero_vq_ema_workspace_bytes(N, K, D)
    ws = _ws(wsb, x_rows.device)
    check(L.pero_vq_ema_accumulate(x_rows.data_ptr(), idx.data_ptr(), N, K, D, out.data_ptr(), ws.data_ptr(), wsb,
                                   _stream()), "pero_vq_ema_accumulate")
    return out


def vq_ema_apply(sums_counts, ema_w, ema_cluster_size, weight, decay, epsilon, codebook=None):
    """In-place EMA update of ema_w / ema_cluster_size / weight (+ refresh of the prepared codebook)."""
    K, D = weight.shape
    for t, n in ((ema_w, "ema_w"), (ema_cluster_size, "ema_cluster_size"), (weight, "weight")):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise TypeError(f"{n} must be a contiguous CUDA float32 tensor")
    ws = _ws(256, weight.device)
    check(_lib.lib().pero_vq_ema_apply(sums_counts.data_ptr(), K, D, float(decay), float(epsilon), ema_w.data_ptr(),
                                       ema_cluster_size.data_ptr(), weight.data_ptr(),
                                       None if codebook is None else codebook.blob.data_ptr(),
                                       0 if codebook is None else codebook.nbytes, ws.data_ptr(), ws.numel(), _stream()),
          "pero_vq_ema_apply")


def kmeans_update(sums_counts, centers, weight_sums, codebook=None):
    """In-place mini-batch k-means centre update (pero_kmeans_update) from the [K*D + K] sums|counts buffer."""
    K, D = centers.shape
    for t, n in ((centers, "centers"), (weight_sums, "weight_sums")):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise TypeError(f"{n} must be a contiguous CUDA float32 tensor")
    check(_lib.lib().pero_kmeans_update(sums_counts.data_ptr(), K, D, centers.data_ptr(), weight_sums.data_ptr(),
                                        None if codebook is None else codebook.blob.data_ptr(),
                                        0 if codebook is None else codebook.nbytes, _stream()), "pero_kmeans_update")


def vq_counts(idx, K):
    counts = torch.empty(int(K), dtype=torch.int64, device=idx.device)
    check(_lib.lib().pero_vq_counts(idx.data_ptr(), idx.numel(), int(K), counts.data_ptr(), _stream()), "pero_vq_counts")
    return counts


# ------------------------------------------------------------------------------------------ MSE
def mse_fwd(a, b, scale_a=1.0, scale_b=0.0):
    """m = mean((a-b)^2); returns the 0-dim tensor scale_a*m + scale_b*m."""
    L = _lib.lib()
    a, b = _f32c(a, "a"), _f32c(b, "b")
    if a.shape != b.shape:
        raise ValueError(f"mse: shapes differ {tuple(a.shape)} vs {tuple(b.shape)}")
    out = torch.empty((), dtype=torch.float32, device=a.device)
    wsb = L.pero_mse_workspace_bytes(a.numel())
    ws = _ws(wsb, a.device)
    check(L.pero_mse_fwd(a.data_ptr(), b.data_ptr(), a.numel(), float(scale_a), float(scale_b), out.data_ptr(), ws.data_ptr(), wsb,
                         _stream()), "pero_mse_fwd")
    return out


def mse_bwd(a, b, coef, grad_out, want_a, want_b):
    """g_b = coef * grad_out * (b - a); g_a = -g_b."""
    a, b = _f32c(a, "a"), _f32c(b, "b")
    g_a = torch.empty_like(a) if want_a else None
    g_b = torch.empty_like(b) if want_b else None
    go = None if grad_out is None else _f32c(grad_out, "grad_out")
    check(_lib.lib().pero_mse_bwd(a.data_ptr(), b.data_ptr(), a.numel(), float(coef), _p(go), _p(g_a), _p(g_b),
                                  _stream()), "pero_mse_bwd")
    return g_a, g_b


def vq_st_commit_bwd(g_quantized, quantized, inputs, coef, grad_loss=None):
    """g_inputs = g_quantized + coef * grad_loss * (inputs - quantized) (same layout for all tensors)."""
    gq, q, x = _f32c(g_quantized, "g_quantized"), _f32c(quantized, "quantized"), _f32c(inputs, "inputs")
    out = torch.empty_like(x)
    gl = None if grad_loss is None else _f32c(grad_loss, "grad_loss")
    check(_lib.lib().pero_vq_st_commit_bwd(gq.data_ptr(), q.data_ptr(), x.data_ptr(), x.numel(), float(coef), _p(gl),
                                           out.data_ptr(), _stream()), "pero_vq_st_commit_bwd")
    return out


# ------------------------------------------------------------------------------------------ masked CE
class PreparedHead:
    """Device blob with the bf16 copy of W and the bias of a Linear(Dh -> V) head."""

    def __init__(self, V, Dh, device):
        self.V, self.Dh = int(V), int(Dh)
        self.nbytes = _lib.lib().pero_head_bytes(self.V, self.Dh)
        self.blob = _ws(self.nbytes, device)

    def prepare(self, W, bias):
        W = _f32c(W, "W")
        b = None if bias is None else _f32c(bias, "bias")
        check(_lib.lib().pero_head_prepare(W.data_ptr(), _p(b), self.V, self.Dh, self.blob.data_ptr(), self.nbytes,
                                           _stream()), "pero_head_prepare")
        return self


def _check_h(h):
    if not h.is_cuda or h.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError(f"hidden states must be CUDA float32 or bfloat16, got {h.dtype}")
    return h if h.is_contiguous() else h.contiguous()


def masked_ce_gather(h, rows, V):
    """Gathers the masked rows of h [N, Dh] into a fresh masked-CE workspace (bf16 GEMM operand + frame -> row map) and
    returns it; reads neither labels nor head, so it can be issued before they exist (masked_ce_fwd(..., ws=...))."""
    L = _lib.lib()
    h = _check_h(h)
    N, Dh = h.shape
    M = rows.numel()
    wsb = L.pero_masked_ce_workspace_bytes(N, M, int(V), Dh)
    ws = _ws(wsb, h.device)
    check(L.pero_masked_ce_gather(h.data_ptr(), 1 if h.dtype == torch.bfloat16 else 0, N, Dh, rows.data_ptr(), M, int(V),
                                  ws.data_ptr(), wsb, _stream()), "pero_masked_ce_gather")
    return ws


def _ce_flags(h, labels_packed, keep_logits=False):
    return (1 if h.dtype == torch.bfloat16 else 0) | (2 if labels_packed else 0) | (4 if keep_logits else 0)


def masked_ce_fwd(h, rows, labels, head, loss_out=None, ws=None, finalize=True, labels_packed=False, keep_logits=False):
    """h [N, Dh]; rows int32 [M]; labels int64 [N].  Returns (loss_sum [1], lse [M], workspace).
    `loss_out`: optional fp32 [1] destination (e.g. a slot of the peer-exchange range next to d_W|d_b).
    `ws`: the workspace masked_ce_gather returned for the same (h, rows): no second gather.
    `finalize=False`: only the logits sweep; (loss_sum, lse) are None and masked_ce_loss(ws, ...) produces them later
    (a backward with ws_from_fwd does not need them).
    `labels_packed`: `labels` are the packed (distance, index) winners of vq_assign(packed=...).
    `keep_logits`: training forward (PERO_CE_KEEP_LOGITS): the softmax numerators of the masked frames (bf16, relative to
    per-chunk maxima) stay in `ws`; pass logits_in_ws=True to the masked_ce_bwd(ws_from_fwd=True) that follows."""
    L = _lib.lib()
    h = _check_h(h)
    N, Dh = h.shape
    M = rows.numel()
    loss_sum = lse = None
    if finalize:
        loss_sum = loss_out if loss_out is not None else torch.empty(1, dtype=torch.float32, device=h.device)
        lse = torch.empty(M, dtype=torch.float32, device=h.device)
    wsb = L.pero_masked_ce_workspace_bytes(N, M, head.V, Dh)
    gathered = ws is not None
    if gathered and ws.numel() < wsb:
        raise ValueError("ws is smaller than pero_masked_ce_workspace_bytes for this shape")
    if not gathered:
        ws = _ws(wsb, h.device)
    check(L.pero_masked_ce_fwd(None if gathered else h.data_ptr(), _ce_flags(h, labels_packed, keep_logits), N, Dh, rows.data_ptr(), M,
                               labels.data_ptr(), head.blob.data_ptr(), head.V, _p(loss_sum), _p(lse),
                               ws.data_ptr(), wsb, _stream()), "pero_masked_ce_fwd")
    return loss_sum, lse, ws


def masked_ce_loss(ws, N, Dh, M, V, loss_out=None):
    """(loss_sum [1], lse [M]) from the log-sum-exp partials a masked_ce_fwd(finalize=False) left in `ws`."""
    L = _lib.lib()
    loss_sum = loss_out if loss_out is not None else torch.empty(1, dtype=torch.float32, device=ws.device)
    lse = torch.empty(int(M), dtype=torch.float32, device=ws.device)
    check(L.pero_masked_ce_loss(int(N), int(Dh), int(M), int(V), loss_sum.data_ptr(), lse.data_ptr(), ws.data_ptr(), ws.numel(),
                                _stream()), "pero_masked_ce_loss")
    return loss_sum, lse


def masked_ce_eval(h, rows, labels, head, ks=(1, 3, 10), want_rank=False):
    """Loss terms and top-k errors of the head on the masked frames, logits never materialised
    (pero_masked_ce_eval).  Returns (loss_sum [1], lse [M], rank int32 [M] or None, errors int64 [len(ks)])."""
    import ctypes
    L = _lib.lib()
    h = _check_h(h)
    N, Dh = h.shape
    M = rows.numel()
    dev = h.device
    loss_sum = torch.empty(1, dtype=torch.float32, device=dev)
    lse = torch.empty(M, dtype=torch.float32, device=dev)
    rank = torch.empty(M, dtype=torch.int32, device=dev) if want_rank else None
    errors = torch.empty(len(ks), dtype=torch.int64, device=dev)
    wsb = L.pero_masked_ce_workspace_bytes(N, M, head.V, Dh)
    ws = _ws(wsb, dev)
    karr = (ctypes.c_int32 * len(ks))(*[int(k) for k in ks])
    check(L.pero_masked_ce_eval(h.data_ptr(), _ce_flags(h, False), N, Dh, rows.data_ptr(), M,
                                labels.data_ptr(), head.blob.data_ptr(), head.V, ctypes.addressof(karr), len(ks),
                                loss_sum.data_ptr(), lse.data_ptr(), _p(rank), errors.data_ptr(), ws.data_ptr(), wsb,
                                _stream()), "pero_masked_ce_eval")
    return loss_sum, lse, rank, errors


def masked_ce_bwd(h, rows, labels, head, lse, grad_scale, inv_count, want_dh=True, ws=None, return_flat=False,
                  want_dw=True, flat_out=None, ws_from_fwd=False, v_range=None, labels_packed=False, want_db=None,
                  dw_out=None, db_out=None, logits_in_ws=False):
    """Returns (d_h [N, Dh] like h or None, d_W [V, Dh] fp32, d_b [V] fp32[, flat buffer holding d_W|d_b]).
    ws_from_fwd: `ws` is the workspace masked_ce_fwd returned for the same (h, rows, labels) and has not been
    touched since: the gathered operands in it are reused instead of gathering again, and the log-sum-exp is rebuilt
    from the forward's partials (`lse` may be None).
    v_range=(v0, v1): only label columns [v0, v1) (multiples of 256 or V): rows v0:v1 of d_W / d_b; needs
    want_dh=False, and a final call with want_dw=False for d_h once every range is done.
    Phases for a data-parallel caller (pero_masked_ce_bwd_range): want_dw only (want_db=False, want_dh=False) -> dlogits +
    d_W, whose exchange can start the moment the GEMM is done; then want_dw=False, want_db=True, want_dh=True -> d_h and
    d_b from the dlogits left in `ws`.  dw_out / db_out: separate destinations (e.g. two peer-exchange ranges).
    logits_in_ws (with ws_from_fwd): the forward ran with keep_logits=True; every phase of this backward must say so."""
    L = _lib.lib()
    h = _check_h(h)
    N, Dh = h.shape
    M = rows.numel()
    if want_db is None:
        want_db = want_dw
    d_h = torch.empty_like(h) if want_dh else None
    # d_W and d_b share one flat buffer so that data-parallel ranks all-reduce them in a single call
    flat = d_W = d_b = None
    if dw_out is not None or db_out is not None:
        d_W = dw_out if want_dw else None
        d_b = db_out if want_db else None
    elif want_dw or want_db:       # neither: second phase, d_h only, from the dlogits a previous call left in `ws`
        flat = flat_out if flat_out is not None else torch.empty(head.V * Dh + head.V, dtype=torch.float32, device=h.device)
        d_W = flat[:head.V * Dh].view(head.V, Dh) if want_dw else None
        d_b = flat[head.V * Dh:] if want_db else None
    wsb = L.pero_masked_ce_workspace_bytes(N, M, head.V, Dh)
    if ws_from_fwd and (ws is None or ws.numel() < wsb):
        raise ValueError("ws_from_fwd needs the forward call's workspace")
    if ws is None or ws.numel() < wsb:
        ws = _ws(wsb, h.device)
    gs = None if grad_scale is None else _f32c(grad_scale, "grad_scale")
    v0, v1 = (0, head.V) if v_range is None else (int(v_range[0]), int(v_range[1]))
    if lse is None and not ws_from_fwd:
        raise ValueError("lse is required unless ws_from_fwd")
    if logits_in_ws and not ws_from_fwd:
        raise ValueError("logits_in_ws needs ws_from_fwd")
    check(L.pero_masked_ce_bwd_range(None if ws_from_fwd else h.data_ptr(), _ce_flags(h, labels_packed, logits_in_ws), N, Dh,
                                     rows.data_ptr(), M, labels.data_ptr(), head.blob.data_ptr(), head.V, _p(lse), _p(gs),
                                     float(inv_count), v0, v1, _p(d_h), _p(d_W), _p(d_b), ws.data_ptr(), wsb,
                                     _stream()), "pero_masked_ce_bwd_range")
    return (d_h, d_W, d_b, flat) if return_flat else (d_h, d_W, d_b)


def head_argmax_codebook(W, bias):
    """The head as a prepared "codebook" of the distance kernel: vq_assign(hidden_rows, blob, ...) then returns
    argmax_v (h.W_v + b_v) per frame (pero_head_argmax_prepare)."""
    W = _f32c(W, "W")
    b = None if bias is None else _f32c(bias, "bias")
    V, Dh = W.shape
    cb = PreparedCodebook(V, Dh, W.device)
    check(_lib.lib().pero_head_argmax_prepare(W.data_ptr(), _p(b), V, Dh, cb.blob.data_ptr(), cb.nbytes, _stream()),
          "pero_head_argmax_prepare")
    return cb


def mask_pixels_(x, rows, tile, frames_per_line):
    """In place: the 8-px column of every masked frame of the line images x [Nl, C, H, W] is overwritten with the noise
    tile [C, H, pw] (TransformerEncoder.mask, models/transformers.py:53-68); rows: int32 masked-frame indices."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4):
        raise TypeError("x must be a contiguous CUDA float32 [Nl, C, H, W] tensor")
    tile = _f32c(tile, "tile")
    Nl, C, H, W = x.shape
    if tuple(tile.shape[:2]) != (C, H):
        raise ValueError(f"tile must be [C={C}, H={H}, patch_width], got {tuple(tile.shape)}")
    if rows.dtype != torch.int32 or not rows.is_cuda:
        raise TypeError("rows must be a CUDA int32 tensor")
    check(_lib.lib().pero_mask_pixels(x.data_ptr(), Nl, C, H, W, rows.data_ptr(), rows.numel(), int(frames_per_line),
                                      tile.shape[2], tile.data_ptr(), _stream()), "pero_mask_pixels")
    return x


_MASK_DTYPES = {torch.int64: 0, torch.int32: 1, torch.uint8: 2, torch.bool: 2}


def mask_compact(mask, labels=None, want=1):
    """Ordered indices of frames with mask == want (and labels >= 0 when given), no host sync.
    Returns (rows int32 [N] (first `count` valid), count int32 [1])."""
    L = _lib.lib()
    m = mask.reshape(-1)
    if not m.is_contiguous():
        m = m.contiguous()
    if m.dtype not in _MASK_DTYPES:
        raise TypeError(f"mask dtype {m.dtype} not supported")
    N = m.numel()
    rows = torch.empty(N, dtype=torch.int32, device=m.device)
    count = torch.zeros(1, dtype=torch.int32, device=m.device)
    wsb = L.pero_mask_compact_workspace_bytes(N)
    ws = _ws(wsb, m.device)
    lab = None
    if labels is not None:       # the kernel reads device int64: coerce instead of reinterpreting other dtypes / CPU memory
        lab = labels.reshape(-1).to(device=m.device, dtype=torch.int64).contiguous()
        if lab.numel() != N:
            raise ValueError("labels and mask must have the same number of elements")
    check(L.pero_mask_compact(m.data_ptr(), _MASK_DTYPES[m.dtype], int(want), _p(lab), N, rows.data_ptr(),
                              count.data_ptr(), ws.data_ptr(), wsb, _stream()), "pero_mask_compact")
    return rows, count


def debug_gemm_tn(a_bf16, b_bf16, variant=0, splits=1):
    ra, kd = a_bf16.shape
    rb = b_bf16.shape[0]
    out = torch.zeros(max(splits, 1), ra, rb, dtype=torch.float32, device=a_bf16.device)
    check(_lib.lib().pero_gemm_tn_bf16(a_bf16.data_ptr(), ra, b_bf16.data_ptr(), rb, kd, int(variant), int(splits),
                                        out.data_ptr(), _stream()), "pero_gemm_tn_bf16")
    return out
