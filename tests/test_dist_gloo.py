"""World-size-2 `gloo` tests (CPU) of the multi-GPU host logic: the collectives, their
reduction ops and the "global count" normalisation are exercised with the oracle standing in for the
device kernels, and must reproduce the single-process result on the concatenated batch."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


# ---------------------------------------------------------------------------------------------- codebook-sharded
def _sharded_assign(rank, world):
    from oracle import pero_oracle as O
    from pero_pretraining_b200.sharding import merge_packed, pack_dist_index_reference, shard_bounds, unpack_index_reference
    g = torch.Generator().manual_seed(5)
    K, D, N = 301, 16, 400
    base = torch.randn(K, D, generator=g)
    base[200] = base[17]                       # exact duplicate across shards: lowest global index must win
    x = torch.randn(N, D, generator=g)
    x[:20] = base[17] + 1e-3 * torch.randn(20, D, generator=g)
    lo, hi = shard_bounds(K, world, rank)
    shard = base[lo:hi]
    d = (shard ** 2).sum(1) - 2 * x @ shard.t()              # what the device epilogue evaluates (|x|^2 dropped)
    dmin, loc = d.min(1)
    loc = torch.argmin(d, dim=1)
    packed = pack_dist_index_reference(d.gather(1, loc[:, None]).squeeze(1), loc + lo)
    merge_packed(packed)                                      # int64 MIN all-reduce
    got = unpack_index_reference(packed)
    full = (base ** 2).sum(1) - 2 * x @ base.t()
    want = torch.argmin(full, dim=1)
    ref64, _, gap = O.assign_fp64(x.numpy(), base.numpy())
    differs = (got.numpy() != ref64)
    return bool(torch.equal(got, want)), bool((got[:20] == 17).all()), bool((gap[differs] < 1e-5).all())


def test_codebook_sharded_min_allreduce_equals_full_argmin():
    for same, dup_ok, ties_ok in _run(_sharded_assign):
        assert same and dup_ok and ties_ok


# ---------------------------------------------------------------------------------------------- batch-sharded EMA
def _dp_ema(rank, world):
    from oracle import pero_oracle as O
    g = torch.Generator().manual_seed(9)
    K, D, nl, T = 24, 8, 6, 10
    w = torch.randn(K, D, generator=g)
    ema_w, cs = w.clone(), torch.ones(K)
    x = torch.randn(nl, D, 1, T, generator=g)
    lo, hi = rank * nl // world, (rank + 1) * nl // world
    mine = O.vq_forward(x[lo:hi], w, ema_w, cs, 0.99, 1e-5, True)
    buf = torch.cat([mine["dw"].reshape(-1), mine["counts"]])         # the [K*D + K] sums|counts buffer
    dist.all_reduce(buf)                                              # SUM over ranks
    dw, counts = buf[:K * D].view(K, D), buf[K * D:]
    new_cs = cs * 0.99 + (1 - 0.99) * counts
    n = new_cs.sum()
    new_cs = (new_cs + 1e-5) / (n + K * 1e-5) * n
    new_ema = ema_w * 0.99 + (1 - 0.99) * dw
    new_w = new_ema / new_cs.unsqueeze(1)
    full = O.vq_forward(x, w, ema_w, cs, 0.99, 1e-5, True)
    ok_idx = torch.equal(mine["indices"], full["indices"].view(nl, T)[lo:hi].reshape(-1))
    return (ok_idx, float((new_w - full["weight"]).abs().max()), float((new_cs - full["ema_cluster_size"]).abs().max()),
            new_w.numpy().tobytes())


def test_dp_ema_allreduce_matches_single_process_and_replicas_stay_identical():
    res = _run(_dp_ema)
    for ok_idx, dw_err, cs_err, _ in res:
        assert ok_idx and dw_err < 1e-5 and cs_err < 1e-6
    assert res[0][3] == res[1][3], "replicated codebooks must stay bit-identical across ranks"


# ---------------------------------------------------------------------------------------------- batch-sharded CE
def _dp_ce(rank, world):
    from oracle import pero_oracle as O
    g = torch.Generator().manual_seed(13)
    Nl, T, Dh, V = 6, 12, 16, 40
    h = torch.randn(Nl, T, Dh, generator=g)
    W, b = torch.randn(V, Dh, generator=g) * 0.2, torch.randn(V, generator=g) * 0.1
    labels = torch.randint(0, V, (Nl, T), generator=g)
    mask = (torch.rand(Nl, T, generator=g) < 0.3).long()
    mask[0:3] = (torch.rand(3, T, generator=g) < 0.6).long()       # unequal M per rank: mean-of-means would be wrong
    lo, hi = rank * Nl // world, (rank + 1) * Nl // world
    hm, lm, mm = h[lo:hi], labels[lo:hi], mask[lo:hi]
    m_local = int(mm.sum())
    loss_local, d_h, d_W, d_b = O.head_masked_ce(hm, W, b, lm, mm)          # local mean and local-mean gradients
    stats = torch.tensor([float(loss_local) * m_local, float(m_local)], dtype=torch.float64)
    dist.all_reduce(stats)                                                  # (loss_sum, M) SUM
    loss = stats[0] / stats[1]
    scale = m_local / float(stats[1])                                       # local-mean grads -> global-mean grads
    flat = torch.cat([(d_W * scale).reshape(-1), d_b * scale])
    dist.all_reduce(flat)                                                   # d_W | d_b SUM
    ref_loss, ref_dh, ref_dW, ref_db = O.head_masked_ce(h, W, b, labels, mask)
    return (abs(float(loss) - float(ref_loss)), float((flat[:V * Dh].view(V, Dh) - ref_dW).abs().max()),
            float((flat[V * Dh:] - ref_db).abs().max()), float((d_h * scale - ref_dh[lo:hi]).abs().max()))


def test_dp_masked_ce_uses_global_count():
    for l_err, w_err, b_err, h_err in _run(_dp_ce):
        assert l_err < 1e-6 and w_err < 1e-6 and b_err < 1e-6 and h_err < 1e-6
