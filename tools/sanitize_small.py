"""Every kernel of libpero_b200.so once, at small shapes, for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck python tools/sanitize_small.py

Small on purpose: the sanitizer slows the GPU code down by one to two orders of magnitude.  The emulated-rank launch
covers the peer-exchange protocol (flag words, slices) on one GPU."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pero_pretraining_b200 import (KMeansLabeller, LinearHead, MiniBatchKMeans, PixelMasker, VectorQuantizer, ops)  # noqa: E402
from pero_pretraining_b200.peer import emulate_all_reduce  # noqa: E402

dev = torch.device("cuda:0")
ops.require_device()
torch.manual_seed(0)
g = torch.Generator().manual_seed(0)

# quantizer forward + EMA (counting sort), commitment loss fwd/bwd, collapsed codebook (long-segment kernels)
K, D, Nl, T = 300, 72, 3, 40
vq = VectorQuantizer(K, D, 0.25, 0.99).to(dev).train()
x = torch.randn(Nl, D, 1, T, generator=g).to(dev).requires_grad_(True)
q, idx = vq(x)
(vq.calculate_loss(q, x) + q.sum()).backward()
xr = torch.randn(500, D, generator=g).to(dev)
ops.vq_ema_accumulate(xr, torch.full((500,), 7, dtype=torch.int64, device=dev), K)        # one long segment
ops.vq_ema_accumulate(torch.randn(9000, 40, generator=g).to(dev), torch.randint(0, 70000, (9000,), generator=g).to(dev), 70000)
ops.vq_ema_accumulate(torch.randn(3000, 40, generator=g).to(dev), torch.randint(0, 20000, (3000,), generator=g).to(dev), 20000)
ops.vq_counts(idx, K)
vq.eval()
vq(x.detach())
# k-means labeller + one mini-batch k-means step; D = 768 streams both operands through the ring
KMeansLabeller(torch.randn(200, 768, generator=g).to(dev)).assign_rows(torch.randn(130, 768, generator=g).to(dev))
MiniBatchKMeans(n_clusters=16, init=np.random.RandomState(0).randn(16, 24).astype(np.float32), device=dev).partial_fit(
    np.random.RandomState(1).randn(300, 24).astype(np.float32))
# codebook-sharded packing
packed = ops.vq_packed_init(130, dev)
cb = ops.PreparedCodebook(200, 64, dev).prepare(torch.randn(200, 64, generator=g).to(dev))
ops.vq_assign(torch.randn(130, 64, generator=g).to(dev), cb, 130, 1, False, index_offset=1000, packed=packed)
ops.vq_unpack(packed, want_dmin=True)
# fused head + masked CE forward / backward / evaluation / argmax, fp32 and bf16 hidden states, wide head
for Dh, V, dt in ((96, 700, torch.float32), (128, 520, torch.bfloat16), (576, 300, torch.float32)):
    head = LinearHead(Dh, V).to(dev)
    h = torch.randn(4, 50, Dh, generator=g).to(dev).to(dt).requires_grad_(True)
    labels = torch.randint(0, V, (4, 50), generator=g).to(dev)
    mask = (np.random.default_rng(0).random((4, 50)) < 0.3).astype(int)
    head.masked_loss(h, labels, mask, 0.5).backward()
    head.masked_errors(h.detach(), labels, mask)
    head.argmax(h.detach())
# logits-in loss, device mask compaction, pixel masking
from pero_pretraining_b200 import MaskedCrossEntropyLoss  # noqa: E402
z = torch.randn(2, 30, 100, generator=g).to(dev).requires_grad_(True)
MaskedCrossEntropyLoss(0.3)(z, torch.randint(0, 100, (2, 30), generator=g).to(dev),
                            (torch.rand(2, 30, generator=g) < 0.4).long().to(dev)).backward()
PixelMasker().to(dev)(torch.rand(2, 3, 40, 100, generator=g).to(dev), (np.random.default_rng(1).random((2, 13)) < 0.5).astype(int))
# 1x1 projections around the quantizer: split-bf16 projection GEMM (channels-first and rows), projected-codebook gather
rows_p, xb_p = ops.proj_forward(torch.randn(3, 24, 37, generator=g).to(dev), torch.randn(16, 24, generator=g).to(dev),
                                torch.randn(16, generator=g).to(dev), 3, 37, True, want_rows=True, want_bf16=True,
                                packed=torch.empty(111, dtype=torch.int64, device=dev))
table_p, _ = ops.proj_forward(torch.randn(50, 70, generator=g).to(dev), torch.randn(300, 70, generator=g).to(dev), None, 50, 1, False)
ops.gather_rows_cf(table_p, torch.randint(0, 50, (111,), generator=g).to(dev), 3, 37)
# peer exchange protocol, 4 emulated ranks in one cooperative launch
bufs = [torch.zeros(16384 + 4096, dtype=torch.uint8, device=dev) for _ in range(4)]
for b in bufs:
    b[16384:].view(torch.float32).fill_(1.0)
emulate_all_reduce(bufs, "sum", 16384, 1024, n_blocks=2)
emulate_all_reduce(bufs, "min", 16384, 512, n_blocks=2)
torch.cuda.synchronize()
print("sanitize_small ok")
