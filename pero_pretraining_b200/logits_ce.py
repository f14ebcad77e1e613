"""MaskedCrossEntropyLoss.forward(output, labels, mask) for callers that already hold the logits
(masked_pretraining/model.py:78-95), running in libpero_b200.so (pero_ce_logits_fwd / _bwd)."""
import torch

from . import _lib, ops
from ._lib import check


class _LogitsCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, labels, terms):
        V = output.shape[-1]
        z = output.detach().reshape(-1, V)
        if z.dtype not in (torch.float32, torch.bfloat16):
            z = z.float()
        z = z.contiguous()
        lab = labels.detach().reshape(-1).long().contiguous()
        L = _lib.lib()
        N = z.shape[0]
        is_bf16 = 1 if z.dtype == torch.bfloat16 else 0
        saved, loss = [], None
        for rows, m, weight in terms:
            if m > 0:
                loss_sum = torch.empty(1, dtype=torch.float32, device=z.device)
                lse = torch.empty(m, dtype=torch.float32, device=z.device)
                ws = ops._ws(4 * m + 256, z.device)
                check(L.pero_ce_logits_fwd(z.data_ptr(), is_bf16, N, V, rows.data_ptr(), m, lab.data_ptr(),
                                           loss_sum.data_ptr(), lse.data_ptr(), ws.data_ptr(), ws.numel(),
                                           ops._stream()), "pero_ce_logits_fwd")
                term = loss_sum[0] / float(m)
            else:
                lse = None
                term = torch.full((), float('nan'), device=z.device)    # F.cross_entropy on an empty selection
            term = term * weight if weight != 1.0 else term
            loss = term if loss is None else loss + term
            saved.append((rows, m, weight, lse))
        ctx.saved, ctx.z, ctx.lab = saved, z, lab
        ctx.out_shape, ctx.out_dtype = output.shape, output.dtype
        return loss

    @staticmethod
    def backward(ctx, g):
        z, lab = ctx.z, ctx.lab
        L = _lib.lib()
        N, V = z.shape
        g = g.detach().float().reshape(1).contiguous()
        d = torch.empty_like(z)
        first = True
        for rows, m, weight, lse in ctx.saved:
            if m == 0:
                continue
            check(L.pero_ce_logits_bwd(z.data_ptr(), 1 if z.dtype == torch.bfloat16 else 0, N, V, rows.data_ptr(), m,
                                       lab.data_ptr(), lse.data_ptr(), g.data_ptr(), weight / float(m), 1 if first else 0,
                                       d.data_ptr(), ops._stream()), "pero_ce_logits_bwd")
            first = False
        if first:
            d.zero_()
        return d.reshape(ctx.out_shape).to(ctx.out_dtype), None, None


def masked_ce_from_logits(output, labels, mask, unmasked_weight=None):
    from .masked_pretraining import _rows_from_mask
    if not output.is_cuda:
        raise _lib.PeroError("MaskedCrossEntropyLoss runs on a CUDA (B200) device only; there is no CPU path")
    rows, m = _rows_from_mask(mask, labels, 1, False, output.device)
    terms = [(rows, m, 1.0)]
    if unmasked_weight is not None:
        rows0, m0 = _rows_from_mask(mask, labels, 0, True, output.device)
        terms.append((rows0, m0, float(unmasked_weight)))
    return _LogitsCE.apply(output, labels, terms)
