"""Drop-in quantizer modules: same names, constructor arguments, attributes, state_dict keys and return
values as the reference's ``pero_pretraining/models/autoencoders.py`` (VectorQuantizer :170-241,
VQVAE :108-167), with the quantize path running in libpero_b200.so.

Differences that are observable only through object identity (documented in DESIGN.md):
  * the EMA step updates ``embedding.weight`` / ``ema_w`` / ``ema_cluster_size`` IN PLACE; the reference
    re-wraps fresh ``torch.nn.Parameter`` objects every step (autoencoders.py:235-237), which silently
    detaches them from any optimizer.  Values and state_dict contents are the same.
  * distances use bf16 operands with fp32 accumulation, so an index may differ from the fp32 reference
    where the two nearest codewords are closer than the epsilon stated in tests/ (near-tie rule).
"""
import torch

from . import ops


class _QuantizeST(torch.autograd.Function):
    """assign -> gather -> straight-through (+ EMA side effect).  Backward is the identity on the inputs
    (autoencoders.py:239: inputs + (quantized - inputs).detach()); nothing flows to the codebook."""

    @staticmethod
    def forward(ctx, inputs, vq):
        n_lines, D = inputs.shape[0], inputs.shape[1]
        frames = 1
        for s in inputs.shape[2:]:
            frames *= s
        x = inputs.detach()
        if x.dtype != torch.float32:
            x = x.float()
        if not x.is_contiguous():
            x = x.contiguous()
        cb = vq._prepared_codebook()
        update = vq.decay > 0.0 and vq.training
        # the peer-memory exchange is graph-replayable (its barrier epochs only grow); a torch.distributed all-reduce
        # inside a captured graph is not attempted
        graphable = vq._use_cuda_graph and x.numel() > 0 and (vq._dp_group is None or vq._peer_range is not None or not update)
        if graphable:
            out, idx = vq._graphed_forward(x, cb, update, n_lines, frames)
        else:
            out, idx = vq._eager_forward(x, cb, update, n_lines, frames)
        if update and (x.numel() > 0 or vq._dp_group is not None):
            vq._codebook_tag = vq._weight_tag()
        ctx.mark_non_differentiable(idx)
        return out.view(inputs.shape), idx

    @staticmethod
    def backward(ctx, g_out, _g_idx):
        return g_out, None


class _ProjectedQuantize(torch.autograd.Function):
    """VQVAE.quantize (models/autoencoders.py:142-147) as one node: encoder 1x1 projection -> assign -> quantize ->
    decoder 1x1 projection (+ EMA side effect).  Forward, all in libpero_b200: the encoder projection is a split-bf16
    tensor-core GEMM whose epilogue writes the distance GEMM's operand directly (the projected NCHW tensor never exists);
    the decoder projection is applied to the K codewords and the output is a row gather of that projected codebook
    (W_d e[idx] + b_d == (E W_d^T + b_d)[idx]).  Backward: the straight-through estimator makes it two dense layers
    back to back, d_tokens -> W_d -> W_e -> d_features; those are plain matrix products (torch.matmul, fp32)."""

    @staticmethod
    def forward(ctx, feats, w_enc, b_enc, w_dec, b_dec, vq, owner=None):
        n_lines, C = feats.shape[0], feats.shape[1]
        frames = 1
        for s in feats.shape[2:]:
            frames *= s
        N = n_lines * frames
        x = feats.detach()
        if x.dtype != torch.float32:
            x = x.float()
        if not x.is_contiguous():
            x = x.contiguous()
        D, Cd = vq.embeddings_dim, w_dec.shape[0]
        we, wd = w_enc.detach().reshape(D, C), w_dec.detach().reshape(Cd, D)
        be = None if b_enc is None else b_enc.detach()
        bd = None if b_dec is None else b_dec.detach()
        cb = vq._prepared_codebook()
        weight = vq.embedding.weight.data
        update = vq.decay > 0.0 and vq.training
        packed = torch.empty(N, dtype=torch.int64, device=x.device)
        # (needs_input_grad reports requires_grad of the parameters also under torch.no_grad(): ask the owner's flag)
        need_grad = any(ctx.needs_input_grad[:5]) and (owner is None or owner._grad_enabled)
        x_rows, xb = ops.proj_forward(x, we, be, n_lines, frames, True, want_rows=update or need_grad, want_bf16=True, packed=packed)
        if N > 0:
            ops.vq_assign_bf16(xb, cb, packed)
        idx, _ = ops.vq_unpack(packed)
        # decoder projection of the (old) codebook, then the gather: the reference quantizes with the weights it had
        # before this step's EMA update (:218-222 come before :225-237)
        table = owner._projected_codebook(weight, w_dec, b_dec) if owner is not None else None
        if table is None:
            table, _ = ops.proj_forward(weight, wd, bd, vq.num_embeddings, 1, False)
            if owner is not None:
                owner._store_projected_codebook(table, weight, w_dec, b_dec)
        tokens = ops.gather_rows_cf(table, idx, n_lines, frames).view((n_lines, Cd) + tuple(feats.shape[2:]))
        if need_grad:
            # the decoder projection's input as the reference's autograd sees it: x + (e[idx] - x)
            q_st = ops.vq_gather_st(x_rows, idx, weight, n_lines, frames, False) if N > 0 else x_rows
            ctx.save_for_backward(x, q_st, we, wd)
        if update and (N > 0 or vq._dp_group is not None):
            vq._ema_update(x_rows, idx, cb)
            vq._codebook_tag = vq._weight_tag()
        ctx.shapes = (feats.shape, w_enc.shape, w_dec.shape, feats.dtype, b_enc is not None, b_dec is not None)
        ctx.mark_non_differentiable(idx)
        return tokens, idx

    @staticmethod
    def backward(ctx, g_tokens, _g_idx):
        x, q_st, we, wd = ctx.saved_tensors
        f_shape, we_shape, wd_shape, f_dtype, has_be, has_bd = ctx.shapes
        n_lines, C = f_shape[0], f_shape[1]
        Cd, D = wd.shape
        g = g_tokens.detach().float().reshape(n_lines, Cd, -1).permute(0, 2, 1).reshape(-1, Cd)        # rows [N, Cd]
        xr = x.reshape(n_lines, C, -1).permute(0, 2, 1).reshape(-1, C)                                 # rows [N, C]
        d_wd = (g.t() @ q_st).view(wd_shape) if ctx.needs_input_grad[3] else None
        d_bd = g.sum(0) if (has_bd and ctx.needs_input_grad[4]) else None
        dq = g @ wd                                             # straight-through: also the projected features' gradient
        d_we = (dq.t() @ xr).view(we_shape) if ctx.needs_input_grad[1] else None
        d_be = dq.sum(0) if (has_be and ctx.needs_input_grad[2]) else None
        d_x = None
        if ctx.needs_input_grad[0]:
            d_x = (dq @ we).view(n_lines, -1, C).permute(0, 2, 1).reshape(f_shape)
            if d_x.dtype != f_dtype:
                d_x = d_x.to(f_dtype)
        return d_x, d_we, d_be, d_wd, d_bd, None, None


class _WeightedMse(torch.autograd.Function):
    """w * mse(tokens, features) with gradient to `features` only (w_features) and/or `tokens` (w_tokens)."""

    @staticmethod
    def forward(ctx, tokens, features, w_tokens, w_features):
        t, f = tokens.detach(), features.detach()
        if t.dtype != torch.float32:
            t = t.float()
        if f.dtype != torch.float32:
            f = f.float()
        if not t.is_contiguous():
            t = t.contiguous()
        if not f.is_contiguous():
            f = f.contiguous()
        ctx.save_for_backward(t, f)
        ctx.w = (float(w_tokens), float(w_features))
        ctx.dtypes = (tokens.dtype, features.dtype)
        # q_latent_loss + commitment_cost * e_latent_loss, each term rounded as the reference does (:200)
        return ops.mse_fwd(t, f, float(w_tokens), float(w_features))

    @staticmethod
    def backward(ctx, g):
        t, f = ctx.saved_tensors
        w_t, w_f = ctx.w
        g = g.detach()
        if g.dtype != torch.float32 or not g.is_contiguous():
            g = g.float().contiguous()
        numel = t.numel()
        g_t = g_f = None
        if ctx.needs_input_grad[0] and w_t != 0.0:
            _, g_t = ops.mse_bwd(t, f, -2.0 * w_t / numel, g, want_a=False, want_b=True)   # w_t*2/n*(t - f)
            if ctx.dtypes[0] != torch.float32:
                g_t = g_t.to(ctx.dtypes[0])
        if ctx.needs_input_grad[1] and w_f != 0.0:
            _, g_f = ops.mse_bwd(t, f, 2.0 * w_f / numel, g, want_a=False, want_b=True)    # w_f*2/n*(f - t)
            if ctx.dtypes[1] != torch.float32:
                g_f = g_f.to(ctx.dtypes[1])
        return g_t, g_f, None, None


class VectorQuantizer(torch.nn.Module):
    """models/autoencoders.py:170-241."""

    def __init__(self, num_embeddings, embeddings_dim, commitment_cost, decay, epsilon=1e-5):
        super().__init__()
        self.embeddings_dim = embeddings_dim
        self.num_embeddings = num_embeddings
        self.embedding = torch.nn.Embedding(self.num_embeddings, self.embeddings_dim)
        if decay > 0.0:
            self.embedding.weight.data.normal_()
            self.register_buffer('ema_cluster_size', torch.zeros(num_embeddings))
            self.ema_w = torch.nn.Parameter(torch.Tensor(num_embeddings, self.embeddings_dim))
            self.ema_w.data.normal_()
        else:
            self.embedding.weight.data.uniform_(-1 / self.num_embeddings, 1 / self.num_embeddings)
        self.commitment_cost = commitment_cost
        self.decay = decay
        self.epsilon = epsilon
        self._codebook = None         # derived cache (bf16 operand, |c|^2): non-persistent, rebuilt on demand
        self._codebook_tag = None
        self._dp_group = None
        self._peer_range = None
        self._use_cuda_graph = False
        self._graphs = {}

    # -- launch-bound small batches: the forward's ~12 kernel launches as one CUDA-graph replay
    def enable_cuda_graph(self, enabled=True):
        """Opt-in: VectorQuantizer.forward (single process, or data parallel with the peer-memory exchange) replays a CUDA
        graph of its kernel sequence, captured once per (input shape, training/eval, state addresses), instead of
        launching ~12 kernels from the host; the input is
        copied into a static buffer and the outputs are copied out of static buffers, so the usual tensor semantics
        hold (nothing returned aliases the graph's buffers).  Worth it when the step is host-bound (8192 frames:
        ~100 us of launches become ~25 us)."""
        self._use_cuda_graph = bool(enabled)
        if not enabled:
            self._graphs = {}
        return self

    def _eager_forward(self, x, cb, update, n_lines, frames):
        """assign -> quantize -> (training, decay > 0) EMA update on contiguous fp32 x [n_lines, D, frames...]."""
        weight = self.embedding.weight.data
        if self._dp_group is None:
            # single process: one call of the C ABI for assign -> quantize -> EMA update
            return ops.vq_forward(x, cb, weight, self.ema_w.data if update else None,
                                  self.ema_cluster_size if update else None, self.decay, self.epsilon, update and x.numel() > 0,
                                  n_lines, frames, channels_first=True)
        # data parallel: the EMA sums|counts are exchanged between accumulate and apply
        idx, _, x_rows = ops.vq_assign(x, cb, n_lines, frames, channels_first=True, want_rows=True)
        out = ops.vq_gather_st(x_rows, idx, weight, n_lines, frames, channels_first=True)
        if update:
            self._ema_update(x_rows, idx, cb)
        return out, idx

    def _ema_update(self, x_rows, idx, cb):
        """EMA codebook update (autoencoders.py:225-237) from the fp32 frame rows and their indices; data parallel: the
        sums|counts are exchanged between accumulate and apply."""
        weight = self.embedding.weight.data
        if self._dp_group is None:
            if idx.numel() == 0:
                return
            sums_counts = ops.vq_ema_accumulate(x_rows, idx, self.num_embeddings)
        # Every rank enters the exchange, also one whose shard of the batch is empty (it contributes zeros): the
        # peers are waiting for it inside their exchange kernel.
        elif self._peer_range is not None:      # [K, D+1] SUM over ranks, in place in the peer-mapped range
            if idx.numel() > 0:
                sums_counts = ops.vq_ema_accumulate(x_rows, idx, self.num_embeddings, out=self._peer_range.tensor)
            else:
                sums_counts = self._peer_range.tensor.zero_()
            self._peer_range.all_reduce_sum_()
        else:
            if idx.numel() > 0:
                sums_counts = ops.vq_ema_accumulate(x_rows, idx, self.num_embeddings)
            else:
                sums_counts = torch.zeros(self.num_embeddings * (self.embeddings_dim + 1), dtype=torch.float32, device=x_rows.device)
            torch.distributed.all_reduce(sums_counts, group=self._dp_group)
        ops.vq_ema_apply(sums_counts, self.ema_w.data, self.ema_cluster_size, weight, self.decay, self.epsilon, cb)

    def _graphed_forward(self, x, cb, update, n_lines, frames):
        weight = self.embedding.weight.data
        key = (tuple(x.shape), bool(update), weight.data_ptr(), cb.blob.data_ptr(),
               self.ema_w.data_ptr() if update else 0, self.ema_cluster_size.data_ptr() if update else 0, x.device,
               self._dp_group is not None)
        g = self._graphs.get(key)
        if g is None:
            g = {"x": torch.empty_like(x)}
            graph = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(torch.cuda.current_stream(x.device))
            # capture records the launches without executing them: the EMA state is not touched here; everything the
            # forward allocates while being captured lives in the graph's private pool and is reused by every replay
            with torch.cuda.graph(graph, stream=side):
                g["out"], g["idx"] = self._eager_forward(g["x"], cb, update, n_lines, frames)
            torch.cuda.current_stream(x.device).wait_stream(side)
            g["graph"] = graph
            if len(self._graphs) >= 8:        # shapes come and go (ragged last batch): keep the cache small
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = g
        g["x"].copy_(x)
        g["graph"].replay()
        return g["out"].clone(), g["idx"].clone()

    # -- data parallel: batch-sharded frames, replicated codebook, EMA sums/counts all-reduced
    def enable_data_parallel(self, group=None, peer=True):
        """peer=True: the EMA sums|counts are exchanged by libpero_b200's own NVLink/NVSwitch kernel through a
        peer-mapped buffer (pero_peer_allreduce_sum_f32); peer=False: torch.distributed.all_reduce (NCCL, or
        gloo in the CPU tests of the semantics)."""
        if not torch.distributed.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self._dp_group = group if group is not None else torch.distributed.group.WORLD
        self._peer_range = None
        if peer:
            from .peer import PeerBuffer, PeerRange
            n = self.num_embeddings * self.embeddings_dim + self.num_embeddings
            buf = PeerBuffer(4 * n + 256, self.embedding.weight.device, self._dp_group)
            self._peer_range = PeerRange(buf, n, torch.float32)
        return self

    def invalidate_codebook(self):
        """Forget the prepared codebook (bf16 operand + |c|^2).  The cache is keyed on (data_ptr, _version) of
        embedding.weight; a write through `.data` (e.g. `vq.embedding.weight.data.copy_(kmeans_centers)`, the idiom
        the reference's own __init__ uses) does not bump `_version`, so call this after one.  load_state_dict() does it
        by itself."""
        self._codebook_tag = None

    def _load_from_state_dict(self, *args, **kwargs):
        self._codebook_tag = None
        return super()._load_from_state_dict(*args, **kwargs)

    def _weight_tag(self):
        w = self.embedding.weight
        return (w.data_ptr(), w._version, w.device)

    def _prepared_codebook(self):
        w = self.embedding.weight
        if not w.is_cuda:
            raise ops._lib.PeroError("VectorQuantizer runs on a CUDA (B200) device only; there is no CPU path")
        if w.dtype != torch.float32 or not w.is_contiguous():
            raise TypeError("embedding.weight must be contiguous float32")
        tag = self._weight_tag()
        if self._codebook is None or self._codebook.blob.device != w.device:
            self._codebook = ops.PreparedCodebook(self.num_embeddings, self.embeddings_dim, w.device)
            self._codebook_tag = None
        if self._codebook_tag != tag:
            self._codebook.prepare(w.data)
            self._codebook_tag = tag
        return self._codebook

    def calculate_loss(self, tokens, features):
        """autoencoders.py:193-202.  Accepts any two same-shape tensors (VQVAE passes tensors from either
        side of its 1x1 projections, :155-159)."""
        w_tokens = 0.0 if self.decay > 0.0 else 1.0
        return _WeightedMse.apply(tokens, features, w_tokens, self.commitment_cost)

    def forward(self, inputs):
        """inputs [Nl, D, H, W] -> (quantized [Nl, D, H, W] with straight-through grad, indices [Nl*H*W] int64)."""
        if inputs.dim() != 4 or inputs.shape[1] != self.embeddings_dim:
            raise ValueError(f"expected [N, {self.embeddings_dim}, H, W], got {tuple(inputs.shape)}")
        return _QuantizeST.apply(inputs, self)


class VQVAE(torch.nn.Module):
    """models/autoencoders.py:108-167; encoder/decoder are the caller's modules (conv stacks are out of scope)."""

    def __init__(self, encoder, decoder, num_embeddings, embeddings_dim, commitment_cost=0.25, decay=0.99,
                 reconstruction_loss='mse'):
        super().__init__()
        self.encoder = encoder
        self.decoder = decoder
        self.encoder_projection_layer = torch.nn.Conv2d(encoder.out_channels, embeddings_dim, 1)
        self.decoder_projection_layer = torch.nn.Conv2d(embeddings_dim, decoder.base_channels, 1)
        self.num_embeddings = num_embeddings
        self.embeddings_dim = embeddings_dim
        self.reconstruction_loss = reconstruction_loss
        self.vq = VectorQuantizer(self.num_embeddings, self.embeddings_dim, commitment_cost, decay)

    def calculate_loss(self, images, reconstructions, features, tokens):
        kind = self.reconstruction_loss.lower()
        if kind in ('l2', 'mse'):
            recon_loss = torch.nn.functional.mse_loss(images, reconstructions)
        elif kind in ('l1', 'mae'):
            recon_loss = torch.nn.functional.l1_loss(images, reconstructions)
        else:
            raise ValueError(f'Unknown reconstruction loss: {self.reconstruction_loss}')
        return self.vq.calculate_loss(tokens, features) + recon_loss

    def encode(self, x):
        return self.encoder(x)

    def decode(self, x):
        return self.decoder(x)

    # 'auto' (default): the fused path where it is the faster one -- label production / evaluation (no gradient: the
    # decoder-projected codebook is computed once per codebook version, the projected NCHW tensor never exists) -- and the
    # torch.nn.Conv2d projections around self.vq in training, where cuDNN's TF32 convolutions beat the fp32-grade
    # three-product GEMM (measured).  True / False force one path.
    fuse_projections = 'auto'
    _grad_enabled = True
    _table = None
    _table_tag = None

    def _table_key(self, weight, w_dec, b_dec):
        return (weight.data_ptr(), weight._version, self.vq._codebook_tag, w_dec.data_ptr(), w_dec._version,
                None if b_dec is None else (b_dec.data_ptr(), b_dec._version))

    def _projected_codebook(self, weight, w_dec, b_dec):
        """decoder_projection_layer applied to the K codewords ([K, Cd] fp32), kept while neither the codebook nor the
        projection has changed (a label-production loop projects the codebook once).  invalidate_projections() after a
        write through `.data`."""
        if self._table is not None and self._table_tag == self._table_key(self.vq.embedding.weight, w_dec, b_dec):
            return self._table
        return None

    def _store_projected_codebook(self, table, weight, w_dec, b_dec):
        if self.training:            # the EMA update behind this forward changes the codebook: nothing to keep
            self._table = self._table_tag = None
            return
        self._table, self._table_tag = table, self._table_key(self.vq.embedding.weight, w_dec, b_dec)

    def invalidate_projections(self):
        self._table = self._table_tag = None
        self.vq.invalidate_codebook()

    def quantize(self, x):
        """autoencoders.py:142-147.  Fused: the two 1x1 projections run inside libpero_b200 around the distance GEMM (see
        _ProjectedQuantize); otherwise they are torch.nn.Conv2d calls around self.vq."""
        enc, dec = self.encoder_projection_layer, self.decoder_projection_layer
        grad = torch.is_grad_enabled() and (x.requires_grad or enc.weight.requires_grad or dec.weight.requires_grad)
        fuse = (not grad) if self.fuse_projections == 'auto' else bool(self.fuse_projections)
        if fuse and x.is_cuda and x.dim() == 4:
            self._grad_enabled = grad
            return _ProjectedQuantize.apply(x, enc.weight, enc.bias, dec.weight, dec.bias, self.vq, self)
        x = enc(x)
        tokens, labels = self.vq(x)
        return dec(tokens), labels

    def labels(self, x):
        """Labels only (what scripts/produce_vqvae_labels.py:37 keeps of quantize()): encoder projection -> nearest
        codeword, [n*h*w] int64.  No quantized output, no decoder projection, no EMA update, no gradient."""
        enc = self.encoder_projection_layer
        if not (x.is_cuda and x.dim() == 4):
            return self.quantize(x)[1]
        with torch.no_grad():
            n_lines, C = x.shape[0], x.shape[1]
            frames = x.shape[2] * x.shape[3]
            N = n_lines * frames
            xf = x.detach()
            if xf.dtype != torch.float32:
                xf = xf.float()
            packed = torch.empty(N, dtype=torch.int64, device=x.device)
            _, xb = ops.proj_forward(xf.contiguous(), enc.weight.detach().reshape(self.embeddings_dim, C),
                                     None if enc.bias is None else enc.bias.detach(), n_lines, frames, True, want_rows=False,
                                     want_bf16=True, packed=packed)
            if N > 0:
                ops.vq_assign_bf16(xb, self.vq._prepared_codebook(), packed)
            return ops.vq_unpack(packed)[0]

    def forward(self, images):
        features = self.encode(images)
        tokens, labels = self.quantize(features)
        reconstructions = self.decode(tokens)
        loss = self.calculate_loss(images, reconstructions, features, tokens)
        return {
            'tokens': tokens,
            'labels': labels,
            'loss': loss,
            'reconstructions': reconstructions,
            'counts': ops.vq_counts(labels, self.num_embeddings),     # torch.bincount(labels, minlength=K), :165
        }
