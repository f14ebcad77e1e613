"""Drop-in head / loss / model glue of the masked-label-prediction task: same class names, constructor
arguments, attributes and state_dict keys as the reference's ``pero_pretraining/masked_pretraining/model.py``
(LinearHead :98-105, MaskedCrossEntropyLoss :72-95, MaskedTransformerEncoder :33-69).

The training path is fused: only the masked frames are pushed through the head, and the [M, V] logits never
reach HBM (C ABI: pero_masked_ce_fwd / pero_masked_ce_bwd).  Full logits for every frame -- which the
reference's evaluation reads from ``result['output']`` (tester.py:70-93, visualizer.py:32) -- are produced
only in eval mode or on request.
"""
import numpy as np
import torch

from . import ops


class _RowStager:
    """Per-device ring of (pinned host buffer, event) slots through which the ordered list of masked frames
    reaches the GPU without a host-blocking copy: a pageable-memory H2D copy would block the host until everything
    already queued on the stream has run.  A slot is reused only once the copy that last used it has executed (its
    event), however far the host has run ahead of the device; nothing is allocated per step."""

    def __init__(self, device, ring=8, capacity=8192):
        self.device = device
        self.host = [torch.empty(capacity, dtype=torch.int32).pin_memory() for _ in range(ring)]
        self.views = [h.numpy() for h in self.host]          # numpy views of the pinned buffers, made once
        self.events = [torch.cuda.Event() for _ in range(ring)]
        self.used = [False] * ring
        self.next = 0

    def stage(self, rows):
        i = self.next
        self.next = (i + 1) % len(self.host)
        n = int(rows.size)
        if self.used[i]:
            self.events[i].synchronize()
        if self.host[i].numel() < n:
            cap = int(n * 1.5) + 16
            self.host[i] = torch.empty(cap, dtype=torch.int32).pin_memory()
            self.views[i] = self.host[i].numpy()
        self.views[i][:n] = rows
        # the device copy is a fresh tensor (it is saved for backward and must not be recycled with the ring)
        out = torch.empty(n, dtype=torch.int32, device=self.device)
        out.copy_(self.host[i].narrow(0, 0, n), non_blocking=True)
        self.events[i].record(torch.cuda.current_stream(self.device))
        self.used[i] = True
        return out


_ROW_STAGERS = {}      # device -> _RowStager (host-side staging buffers only; no model state)


def _pinned_rows(rows, device):
    st = _ROW_STAGERS.get(device)
    if st is None:
        st = _ROW_STAGERS[device] = _RowStager(device)
    return st.stage(rows)


def create_mask(labels, masking_prob, rng=None):
    """BatchOperator._create_mask (masked_pretraining/batch_operator.py:27-32): mask = (rand < p) * (labels >= 0) as a
    host numpy int array, so that the masked-frame list and its length M are known without a device sync.
    rng=None draws from numpy's global generator exactly like the reference (same seed -> same mask); a
    numpy Generator / RandomState may be passed instead."""
    labels = np.asarray(labels)
    active = (labels >= 0).astype(int)
    if rng is None:
        draw = np.random.rand(*labels.shape)
    elif hasattr(rng, "random"):
        draw = rng.random(labels.shape)
    else:
        draw = rng.rand(*labels.shape)
    return (draw < masking_prob).astype(int) * active


def _rows_from_mask(mask, labels, want, require_label, device):
    """Ordered int32 frame indices with mask == want.  A numpy mask (what BatchOperator._create_mask
    returns, batch_operator.py:27-32) is compacted on the host: M is known without a device sync.  A tensor
    mask is compacted on the device and M is read back once (the reference's boolean indexing syncs too)."""
    if isinstance(mask, np.ndarray):
        sel = mask.reshape(-1) == want
        if require_label:
            lab = labels.detach().cpu().numpy().reshape(-1) if isinstance(labels, torch.Tensor) else np.asarray(labels).reshape(-1)
            sel = sel & (lab >= 0)
        rows = np.flatnonzero(sel).astype(np.int32)
        if device.type != "cuda":
            return torch.from_numpy(rows), int(rows.size)
        return _pinned_rows(rows, device), int(rows.size)
    mask = mask.to(device)
    rows, count = ops.mask_compact(mask, labels if require_label else None, want)
    m = int(count.item())
    return rows[:m], m


def masked_rows(mask, device, labels=None):
    """(device int32 list of the frames with mask == 1 in ascending order, its length M) for a numpy or tensor mask: the
    one trip of the mask to the GPU that both the pixel masking and the masked cross-entropy use."""
    return _rows_from_mask(mask, labels, 1, False, torch.device(device))


class PixelMasker(torch.nn.Module):
    """TransformerEncoder.mask (models/transformers.py:27-34, 53-68) on the device: the 8-px image column of every masked
    frame is overwritten, in place, with the reference's fixed noise tile (np.random.seed(42); np.random.rand(1, C,
    patch_h, patch_w), reproduced here with a private RandomState so that numpy's global generator is not reseeded)."""

    def __init__(self, height=40, patch_size=(40, 8), in_channels=3):
        super().__init__()
        self.height, self.patch_size, self.in_channels = height, patch_size, in_channels
        tile = np.random.RandomState(42).rand(1, in_channels, patch_size[0], patch_size[1])
        self.register_buffer("mask_tile", torch.tensor(tile[0], dtype=torch.float32), persistent=False)

    def forward(self, x, mask=None, rows=None):
        """x [N, C, H, W] float32 (modified in place and returned, like the reference); mask [N, W/8] {0,1} numpy/tensor,
        or rows = masked_rows(mask, x.device) when the caller has staged the list already."""
        if rows is None:
            if mask is None:
                return x
            rows = masked_rows(mask, x.device)
        r, m = rows
        if m == 0:
            return x
        pw = self.patch_size[1]
        frames = (x.shape[3] + pw - 1) // pw
        return ops.mask_pixels_(x, r, self.mask_tile, frames)


class _FusedHeadCE(torch.autograd.Function):
    """sum-of-terms masked CE of Linear(h) against labels; each term = (rows, weight)."""

    @staticmethod
    def forward(ctx, h, W, b, labels, head_prep, terms, dp_group, peer_range=None):
        # terms: list of (rows int32 tensor, M_local int, weight float)
        n_frames = h.shape[0] * h.shape[1] if h.dim() == 3 else h.shape[0]
        h2 = h.detach().reshape(n_frames, h.shape[-1])
        if h2.dtype not in (torch.float32, torch.bfloat16):
            h2 = h2.float()
        if not h2.is_contiguous():
            h2 = h2.contiguous()
        # the kernels read device int64 labels: coerce other dtypes / host tensors instead of reinterpreting them
        lab = labels.detach().reshape(-1).to(device=h2.device, dtype=torch.int64)
        if not lab.is_contiguous():
            lab = lab.contiguous()
        saved, loss = [], None
        # a forward that will be differentiated leaves the softmax numerators of the masked frames in its workspace (bf16,
        # relative to per-chunk maxima); the backward scales them in place instead of running the logits GEMM again
        keep = bool(ctx.needs_input_grad[0] or ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
        ctx.keep_logits = keep
        for rows, m_local, weight in terms:
            if m_local > 0:
                loss_sum, lse, ws = ops.masked_ce_fwd(h2, rows, lab, head_prep, keep_logits=keep)
            else:
                loss_sum, lse, ws = torch.zeros(1, device=h2.device), None, None
            if dp_group is not None:
                # (loss_sum, M) summed over the ranks in ONE small exchange; the global count stays on the device
                # (scale = weight / M_global is applied by the kernels through their device-scalar argument), so the
                # step has no host synchronisation here.  Empty global selection: 0 * (w / 0) = NaN, like the reference.
                if peer_range is not None:          # the library's own peer kernel on a 16-byte range of the peer buffer
                    stats_range = peer_range.stats[len(saved)]
                    stats = stats_range.tensor
                    stats[0:1].copy_(loss_sum)
                    stats[1].fill_(float(m_local))
                    stats_range.all_reduce_sum_(n_blocks=1)
                    stats = stats.clone()           # the 16-byte slot is reused by the next step
                else:
                    stats = torch.empty(2, dtype=torch.float32, device=h2.device)
                    stats[0:1].copy_(loss_sum)
                    stats[1].fill_(float(m_local))
                    torch.distributed.all_reduce(stats, group=dp_group)
                scale = weight / stats[1:2]                                  # [1] device tensor
                term = (stats[0:1] * scale).view(())
            else:
                m_global = float(m_local)
                scale = weight / m_global if m_global > 0 else float('nan')   # host float
                # mean over the number of selected frames; an empty selection gives NaN like F.cross_entropy
                term = loss_sum.view(()) * scale
            loss = term if loss is None else loss + term
            saved.append((rows, m_local, scale, lse, ws))
        ctx.saved = saved
        ctx.h2, ctx.lab, ctx.head_prep, ctx.dp_group = h2, lab, head_prep, dp_group
        ctx.peer_range = peer_range
        ctx.h_shape, ctx.h_dtype, ctx.has_bias = h.shape, h.dtype, b is not None
        ctx.w_dtype = W.dtype
        return loss

    @staticmethod
    def backward(ctx, g):
        h2, lab, head = ctx.h2, ctx.lab, ctx.head_prep
        g = g.detach().reshape(1)
        if g.dtype != torch.float32:
            g = g.float()
        want_dh = ctx.needs_input_grad[0]
        d_h = d_W = d_b = None
        peer = ctx.peer_range if ctx.dp_group is not None else None
        flat = None
        for rows, m_local, scale, lse, ws in ctx.saved:
            if m_local == 0:
                continue
            first = flat is None
            if isinstance(scale, torch.Tensor):          # data parallel: device scalar weight / M_global
                gs, inv = g * scale, 1.0
            else:
                gs, inv = g, scale
            # the forward's workspace still holds the gathered operands of this term: no second gather.  The in-place
            # conversion of the kept logits consumes them: a second backward through the same graph (retain_graph)
            # recomputes from h and the saved log-sum-exp instead.
            fresh = not getattr(ctx, "ws_consumed", False)
            dh_t, _, _, flat_t = ops.masked_ce_bwd(h2, rows, lab, head, lse, gs, inv, want_dh=want_dh,
                                                   return_flat=True, ws=ws if fresh else None, ws_from_fwd=fresh,
                                                   logits_in_ws=ctx.keep_logits and fresh,
                                                   flat_out=peer.tensor if (first and peer is not None) else None)
            d_h = dh_t if d_h is None else (d_h + dh_t if dh_t is not None else d_h)
            if first:
                flat = flat_t
            else:
                flat.add_(flat_t)
        ctx.ws_consumed = True
        if flat is None:
            flat = peer.tensor.zero_() if peer is not None else torch.zeros(head.V * head.Dh + head.V, device=h2.device)
            if want_dh:
                d_h = torch.zeros_like(h2)
        if ctx.dp_group is not None:
            # every rank ends with the gradient of the single-process loss on the concatenated batch
            if peer is not None:
                peer.all_reduce_sum_()
                if not getattr(peer, "alias_grads", False):
                    flat = flat.clone()        # the exchange range is reused by the next step
            else:
                torch.distributed.all_reduce(flat, group=ctx.dp_group)
        d_W, d_b = flat[:head.V * head.Dh].view(head.V, head.Dh), flat[head.V * head.Dh:]
        if d_h is not None:
            d_h = d_h.view(ctx.h_shape)
            if d_h.dtype != ctx.h_dtype:
                d_h = d_h.to(ctx.h_dtype)
        if d_W.dtype != ctx.w_dtype:
            d_W = d_W.to(ctx.w_dtype)
        return d_h, d_W, (d_b if ctx.has_bias else None), None, None, None, None, None


class LinearHead(torch.nn.Module):
    """masked_pretraining/model.py:98-105.  ``forward`` returns logits for every frame (a plain library GEMM,
    used by evaluation); training goes through ``masked_loss`` instead."""

    def __init__(self, in_features=512, out_features=4096):
        super().__init__()
        self.linear = torch.nn.Linear(in_features, out_features)
        self._prep = None
        self._prep_tag = None
        self._peer_range = None

    def forward(self, x):
        return self.linear(x)

    def enable_peer_exchange(self, group=None, alias_grads=False):
        """Data-parallel gradients d_W | d_b (and the (loss_sum, M) pair of the forward) of masked_loss(dp_group=...) are
        then reduced over the ranks by libpero_b200's own NVLink/NVSwitch kernel in a peer-mapped buffer instead of
        torch.distributed.
        alias_grads=True: the gradients handed to autograd are VIEWS of the exchange range (no 16.8 MB copy per step at
        the bench shape).  The range is overwritten by the next backward, so this is only valid when every backward is
        followed by the optimizer step (no gradient accumulation over several backward passes)."""
        from .peer import PeerBuffer, PeerRange
        W = self.linear.weight
        n = W.shape[0] * W.shape[1] + W.shape[0]
        buf = PeerBuffer(4 * n + 1024, W.device, group)
        self._peer_range = PeerRange(buf, n, torch.float32)
        self._peer_range.alias_grads = bool(alias_grads)
        self._peer_range.stats = [PeerRange(buf, 4, torch.float32) for _ in range(2)]     # one per loss term
        return self

    def invalidate(self):
        """Forget the prepared bf16 operands.  The caches are keyed on (data_ptr, _version) of weight and bias; a write
        through `.data` (e.g. `head.linear.weight.data.copy_(...)`) does not bump `_version`, so call this after one.
        load_state_dict() does it by itself."""
        self._prep_tag = None
        self._argmax_tag = None

    def _load_from_state_dict(self, *args, **kwargs):
        self._prep_tag = None
        self._argmax_tag = None
        return super()._load_from_state_dict(*args, **kwargs)

    def _prepared(self):
        W, b = self.linear.weight, self.linear.bias
        if not W.is_cuda:
            raise ops._lib.PeroError("LinearHead.masked_loss runs on a CUDA (B200) device only; there is no CPU path")
        tag = (W.data_ptr(), W._version, None if b is None else (b.data_ptr(), b._version), W.device)
        if self._prep is None or self._prep.blob.device != W.device:
            self._prep = ops.PreparedHead(W.shape[0], W.shape[1], W.device)
            self._prep_tag = None
        if self._prep_tag != tag:
            self._prep.prepare(W.detach().float(), None if b is None else b.detach().float())
            self._prep_tag = tag
        return self._prep

    def masked_loss(self, hidden, labels, mask, unmasked_weight=None, dp_group=None, rows=None):
        """Fused LinearHead + MaskedCrossEntropyLoss on hidden states [Nl, T, Dh] (or [N, Dh]).
        Any hidden size that is a multiple of 4 takes the fused path (up to 512 the masked rows stay resident in shared
        memory); other sizes go through the logits-in kernels on `self.linear(hidden)` (single process only)."""
        dev = hidden.device
        if hidden.shape[-1] % 4 != 0:
            if dp_group is not None:
                raise ops._lib.PeroError("data-parallel masked_loss needs a hidden size that is a multiple of 4")
            from .logits_ce import masked_ce_from_logits
            return masked_ce_from_logits(self.linear(hidden), labels, mask, unmasked_weight)
        # rows = (device int32 list of the frames with mask == 1, its length): what masked_rows(mask, device) returns,
        # when the caller has staged it already (MaskedTransformerEncoder shares it with the pixel masking)
        rows, m = rows if rows is not None else _rows_from_mask(mask, labels, 1, False, dev)
        terms = [(rows, m, 1.0)]
        if unmasked_weight is not None:
            rows0, m0 = _rows_from_mask(mask, labels, 0, True, dev)     # model.py:85-90
            terms.append((rows0, m0, float(unmasked_weight)))
        return _FusedHeadCE.apply(hidden, self.linear.weight, self.linear.bias, labels, self._prepared(), terms, dp_group,
                                  self._peer_range if dp_group is not None else None)


    def argmax(self, hidden):
        """torch.argmax(self(hidden), dim=-1) for every frame -- what MaskedVisualizer shows as predictions
        (masked_pretraining/visualizer.py:32) -- without materialising the [N, V] logits: the head is handed to the
        distance kernel as a codebook, arg-min of -2 (h.W_v + b_v).  Lowest label on exact ties, like torch.argmax.
        bf16 operands: the winner may differ from the fp32 argmax where the two best logits are within bf16 rounding."""
        W, b = self.linear.weight, self.linear.bias
        if not W.is_cuda:
            raise ops._lib.PeroError("LinearHead.argmax runs on a CUDA (B200) device only; there is no CPU path")
        tag = (W.data_ptr(), W._version, None if b is None else (b.data_ptr(), b._version), W.device)
        if getattr(self, "_argmax_tag", None) != tag:
            self._argmax_cb = ops.head_argmax_codebook(W.detach().float(), None if b is None else b.detach().float())
            self._argmax_tag = tag
        h2 = hidden.detach().reshape(-1, hidden.shape[-1]).float().contiguous()
        idx, _, _ = ops.vq_assign(h2, self._argmax_cb, h2.shape[0], 1, channels_first=False)
        return idx.view(hidden.shape[:-1])

    def masked_errors(self, hidden, labels, mask, ks=(1, 3, 10)):
        """Evaluation on the masked frames without materialising logits: what Tester.test_step + _update_errors
        compute from result['output'] (masked_pretraining/tester.py:57-93).  Returns a dict with the mean masked loss
        (0-dim tensor), 'length' (number of masked frames) and 'errors_<k>' (int tensors, still on the device: a
        caller accumulates them over batches and reads them back once)."""
        dev = hidden.device
        rows, m = _rows_from_mask(mask, labels, 1, False, dev)
        out = {'length': m}
        if m == 0:
            out['loss'] = torch.full((), float('nan'), device=dev)
            for k in ks:
                out[f'errors_{k}'] = torch.zeros((), dtype=torch.int64, device=dev)
            return out
        h2 = hidden.detach().reshape(-1, hidden.shape[-1])
        if h2.dtype not in (torch.float32, torch.bfloat16):
            h2 = h2.float()
        lab = labels.detach().reshape(-1).long().contiguous()
        loss_sum, _, _, errors = ops.masked_ce_eval(h2.contiguous(), rows, lab, self._prepared(), ks)
        out['loss'] = loss_sum.view(()) / m
        for i, k in enumerate(ks):
            out[f'errors_{k}'] = errors[i]
        return out


def update_errors(errors, result):
    """Accumulate LinearHead.masked_errors results the way Tester._update_errors does (tester.py:70-93):
    errors['errors_<k>'] += count, errors['length'] += number of masked frames.  Counts stay device tensors."""
    for key, value in result.items():
        if key.startswith('errors_') or key == 'length':
            errors[key] = errors.get(key, 0) + value
    return errors


class MaskedCrossEntropyLoss(torch.nn.Module):
    """masked_pretraining/model.py:72-95: logits-in interface kept for callers that already hold logits."""

    def __init__(self, unmasked_weight=None):
        super().__init__()
        self.unmasked_weight = unmasked_weight

    def forward(self, output, labels, mask):
        from .logits_ce import masked_ce_from_logits
        return masked_ce_from_logits(output, labels, mask, self.unmasked_weight)


class MaskedTransformerEncoder(torch.nn.Module):
    """masked_pretraining/model.py:33-69.  `backbone(images, mask=mask)` -> [n, c, w] is the caller's module."""

    def __init__(self, backbone, head, loss=None, output='auto', pixel_masker=None):
        """pixel_masker: a PixelMasker -> the input masking of the backbone (models/transformers.py:53-68) runs here, on the
        device, from the same staged masked-frame list as the loss, and the backbone is called with mask=None (its own
        mask() would redo the same overwrite after two more host->device trips of the mask)."""
        super().__init__()
        self.backbone = backbone
        self.head = head
        self.loss = MaskedCrossEntropyLoss() if loss is None else loss
        self.output_mode = output      # 'auto': logits for every frame only in eval mode; True / False to force
        self.pixel_masker = pixel_masker
        self._dp_group = None

    def enable_data_parallel(self, group=None, peer=True):
        if not torch.distributed.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self._dp_group = group if group is not None else torch.distributed.group.WORLD
        if peer and isinstance(self.head, LinearHead):
            self.head.enable_peer_exchange(self._dp_group)
        return self

    def hidden(self, images, mask=None):
        x = self.backbone(images, mask=mask)
        return x.permute(0, 2, 1)          # 'n c w -> n w c' (model.py:60)

    def encode(self, images, mask=None):
        return self.head(self.hidden(images, mask))

    def predict(self, x, mask=None):
        """Label prediction of every frame, int64 [n, w]: torch.argmax(forward(...)['output'], dim=-1) of
        masked_pretraining/visualizer.py:32 without the [n, w, V] logits."""
        hidden = self.hidden(x, mask)
        if isinstance(self.head, LinearHead):
            return self.head.argmax(hidden)
        return torch.argmax(self.head(hidden), dim=-1)

    def forward(self, x, labels=None, mask=None):
        rows = None
        if mask is not None and self.pixel_masker is not None:
            rows = masked_rows(mask, x.device)
            x = self.pixel_masker(x, rows=rows)
            hidden = self.hidden(x, None)
        else:
            hidden = self.hidden(x, mask)
        want_output = (not self.training) if self.output_mode == 'auto' else bool(self.output_mode)
        output = self.head(hidden) if want_output else None
        loss = None
        if mask is not None and labels is not None:
            fused = isinstance(self.head, LinearHead) and isinstance(self.loss, MaskedCrossEntropyLoss)
            if fused:
                loss = self.head.masked_loss(hidden.contiguous(), labels, mask, self.loss.unmasked_weight, self._dp_group,
                                             rows=rows)
            else:
                if output is None:
                    output = self.head(hidden)
                if not isinstance(mask, torch.Tensor):
                    mask = torch.from_numpy(mask).to(output.device)
                loss = self.loss(output, labels, mask)
        return {'output': output, 'loss': loss}

    def save(self, path):
        torch.save(self.state_dict(), path)

    def load(self, path):
        self.load_state_dict(torch.load(path))
