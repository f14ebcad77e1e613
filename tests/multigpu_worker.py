"""Multi-GPU checks of the sharded modes, run under torchrun with one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu_worker.py

  1. peer all-reduce (SUM f32 / MIN i64), multimem and peer-pointer transports: equal to the rank-ordered
     reference, bit-identical on every rank, replayable (also inside a CUDA graph);
  2. batch-sharded VectorQuantizer (3 EMA steps) == the oracle's single-process run on the concatenated batch,
     replicas bit-identical;
  3. batch-sharded LinearHead.masked_loss == the oracle on the concatenated batch (global masked count);
  4. codebook-sharded assign == assign against the full codebook on one GPU.
Prints "MULTIGPU OK" from rank 0 when everything holds; any failure raises (non-zero exit).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

EPS_TIE = 2e-3


def log(msg):
    if dist.get_rank() == 0:
        print(msg, flush=True)


def gather_bytes(t):
    """All ranks' copies of `t` as a list of CPU tensors (on every rank)."""
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t.contiguous())
    return [o.cpu() for o in out]


def check_peer_allreduce(dev, rank, world):
    from pero_pretraining_b200.peer import PeerBuffer, PeerRange
    for use_mc in (True, False):
        n = 8192 * 256 + 8192
        buf = PeerBuffer(4 * n + 8 * 5000 + 1024, dev, use_multicast=use_mc)
        rng_f = PeerRange(buf, n, torch.float32)
        rng_i = PeerRange(buf, 4999, torch.int64)
        g = torch.Generator().manual_seed(77)
        vals = [torch.randn(n, generator=g) * (10.0 ** r) for r in range(world)]
        ints = [torch.randint(-2 ** 62, 2 ** 62, (4999,), generator=g, dtype=torch.int64) for _ in range(world)]
        want = vals[0].clone()
        for r in range(1, world):
            want = want + vals[r]
        want_min = torch.stack(ints).min(0).values
        first_bits = None
        stable = True
        for rep in range(6):
            rng_f.tensor.copy_(vals[rank])
            rng_i.tensor.copy_(ints[rank])
            rng_f.all_reduce_sum_()
            rng_i.all_reduce_min_()
            torch.cuda.synchronize()
            got = rng_f.tensor.cpu()
            if buf.multicast:       # switch-side reduction: order of the adds is the switch's, not rank order
                assert torch.allclose(got, want, rtol=1e-5, atol=1e-5 * float(want.abs().max())), f"multimem sum rep {rep}"
            else:
                assert torch.equal(got, want), f"peer sum rep {rep}"
            assert torch.equal(rng_i.tensor.cpu(), want_min), f"min rep {rep}"
            copies = gather_bytes(rng_f.tensor)
            assert all(torch.equal(copies[0], c) for c in copies[1:]), "replicas differ"
            # run-to-run determinism of the SUM (north_star: "deterministic"): the peer-pointer transport adds in rank
            # order and must repeat bit for bit; for the switch-side reduction PTX leaves the order to the switch, so
            # the observation is reported (DESIGN.md section 5 states the outcome and the deterministic mode)
            if first_bits is None:
                first_bits = got.clone()
            elif not torch.equal(first_bits, got):
                stable = False
        if buf.multicast:
            log(f"  multimem SUM bit-identical over 6 repeats: {stable}")
        else:
            assert stable, "peer-pointer SUM must be bit-identical run to run"
        # CUDA-graph replay of produce -> exchange
        src = vals[rank].to(dev)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            rng_f.tensor.copy_(src)
            rng_f.all_reduce_sum_()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            rng_f.tensor.copy_(src)
            rng_f.all_reduce_sum_()
        for _ in range(5):
            graph.replay()
        torch.cuda.synchronize()
        got = rng_f.tensor.cpu()
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-5 * float(want.abs().max())), "graph replay"
        # timing of the 8.4 MB EMA exchange and a 16.8 MB gradient exchange (device events, max over ranks)
        big = PeerBuffer(4 * (8192 * 512 + 8192) + 2048, dev, use_multicast=use_mc)
        rng_g = PeerRange(big, 8192 * 512 + 8192, torch.float32)
        tiny = PeerRange(big, 4, torch.float32)
        for name, rg in (("16 B (fixed cost)", tiny), ("ema 8.4 MB", rng_f), ("grad 16.8 MB", rng_g)):
            for blocks in (8, 16, 24, 32, 48):
                rg.tensor.zero_()
                for _ in range(3):
                    rg.all_reduce_sum_(blocks)
                torch.cuda.synchronize()
                dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20):
                    rg.all_reduce_sum_(blocks)
                e1.record()
                torch.cuda.synchronize()
                t = torch.tensor([e0.elapsed_time(e1) / 20 * 1e3], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                log(f"  [{buf.transport}] {name} all-reduce, {blocks} CTAs: {float(t):.1f} us")
        for name, numel in (("nccl ema 8.4 MB", n), ("nccl grad 16.8 MB", 8192 * 512 + 8192)):
            x = torch.zeros(numel, device=dev)
            for _ in range(3):
                dist.all_reduce(x)
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                dist.all_reduce(x)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / 20 * 1e3], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            log(f"  {name} all-reduce: {float(t):.1f} us")
        log(f"peer all-reduce OK ({buf.transport}; multicast requested={use_mc})")
        if not buf.multicast and use_mc:
            break               # no multicast object on this box: the second pass would repeat the same transport


def check_dp_quantizer(dev, rank, world):
    from oracle import pero_oracle as O
    from pero_pretraining_b200 import VectorQuantizer
    K, D, nl, T = 512, 64, 8 * world, 32
    g = torch.Generator().manual_seed(21)
    w0 = torch.randn(K, D, generator=g)
    steps = [w0[torch.randint(0, K, (nl * T,), generator=g)].view(nl, T, D).permute(0, 2, 1).reshape(nl, D, 1, T)
             + 0.3 * torch.randn(nl, D, 1, T, generator=g) for _ in range(3)]
    for peer, graphed in ((True, False), (True, True), (False, False)):
        vq = VectorQuantizer(K, D, 0.25, 0.99).to(dev).train()
        with torch.no_grad():
            vq.embedding.weight.copy_(w0); vq.ema_w.copy_(w0); vq.ema_cluster_size.fill_(1.0)
        vq.enable_data_parallel(peer=peer)
        vq.enable_cuda_graph(graphed)          # the exchange kernel replays inside the captured forward
        ref = dict(weight=w0.clone(), ema_w=w0.clone(), cs=torch.ones(K))
        lo, hi = rank * nl // world, (rank + 1) * nl // world
        for x in steps:
            q, idx = vq(x[lo:hi].to(dev))
            out = O.vq_forward(x, ref["weight"], ref["ema_w"], ref["cs"], 0.99, 1e-5, True)
            want_idx = out["indices"].view(nl, T)[lo:hi].reshape(-1)
            bad = (idx.cpu() != want_idx).nonzero().flatten()
            if bad.numel():
                _, _, gap = O.assign_fp64(O.flatten_frames(x[lo:hi])[0].numpy(), ref["weight"].numpy())
                assert (gap[bad.numpy()] < EPS_TIE).all(), "index differs outside a near-tie"
            ref.update(weight=out["weight"], ema_w=out["ema_w"], cs=out["ema_cluster_size"])
            if bad.numel() == 0:
                assert torch.allclose(vq.embedding.weight.detach().cpu(), ref["weight"], rtol=2e-5, atol=2e-6)
                assert torch.allclose(vq.ema_cluster_size.cpu(), ref["cs"], rtol=1e-5, atol=1e-6)
        copies = gather_bytes(vq.embedding.weight.detach())
        assert all(torch.equal(copies[0], c) for c in copies[1:]), "replicated codebooks diverged"
        log(f"data-parallel VectorQuantizer OK (peer={peer}, cuda graph={graphed})")


def check_dp_head(dev, rank, world):
    from oracle import pero_oracle as O
    from pero_pretraining_b200 import LinearHead
    Nl, T, Dh, V = 4 * world, 64, 128, 512
    g = torch.Generator().manual_seed(33)
    h = torch.randn(Nl, T, Dh, generator=g)
    labels = torch.randint(0, V, (Nl, T), generator=g)
    rng = np.random.default_rng(5)
    mask = (rng.random((Nl, T)) < 0.15).astype(int)
    mask[:2] = (rng.random((2, T)) < 0.5).astype(int)          # unequal masked counts per rank
    lo, hi = rank * Nl // world, (rank + 1) * Nl // world
    for peer in (True, False):
        torch.manual_seed(3)
        head = LinearHead(Dh, V).to(dev)
        if peer:
            head.enable_peer_exchange()
        W, b = head.linear.weight.detach().cpu(), head.linear.bias.detach().cpu()
        hm = h[lo:hi].to(dev).requires_grad_(True)
        loss = head.masked_loss(hm, labels[lo:hi].to(dev), mask[lo:hi], None, dist.group.WORLD)
        loss.backward()
        ref_loss, ref_dh, ref_dW, ref_db = O.head_masked_ce(h, W, b, labels, torch.from_numpy(mask))
        assert abs(float(loss) - float(ref_loss)) <= 3e-3 * abs(float(ref_loss)), (float(loss), float(ref_loss))
        for got, want, nm in ((head.linear.weight.grad.cpu(), ref_dW, "d_W"), (head.linear.bias.grad.cpu(), ref_db, "d_b"),
                              (hm.grad.cpu(), ref_dh[lo:hi], "d_h")):
            assert float((got - want).abs().max()) <= 2e-2 * float(want.abs().max()), nm
        copies = gather_bytes(head.linear.weight.grad)
        assert all(torch.equal(copies[0], c) for c in copies[1:]), "replicated gradients differ"
        log(f"data-parallel LinearHead.masked_loss OK (peer={peer})")


def check_sharded_codebook(dev, rank, world):
    from pero_pretraining_b200 import ShardedCodebook, ops
    K, D, N = 4096 + 6, 128, 5000
    g = torch.Generator().manual_seed(41)
    w = torch.randn(K, D, generator=g)
    w[K - 3] = w[11]                                 # exact duplicate in another shard: lowest global index wins
    x = w[torch.randint(0, K, (N,), generator=g)] + 0.4 * torch.randn(N, D, generator=g)
    x[:7] = w[11]
    wd, xd = w.to(dev), x.to(dev)
    full = ops.PreparedCodebook(K, D, dev).prepare(wd)
    want, want_d, _ = ops.vq_assign(xd, full, N, 1, False, want_dmin=True)
    for peer_frames in (N, 0):
        sc = ShardedCodebook(wd, K, rank, world, peer_frames=peer_frames)
        for _ in range(2):
            idx, dmin = sc.assign(xd, N, 1, False, want_dmin=True)
            assert torch.equal(idx, want), f"sharded assign differs (peer_frames={peer_frames})"
            assert torch.equal(dmin, want_d)
        assert bool((idx[:7] == 11).all())
        log(f"codebook-sharded assign OK (peer={'yes' if peer_frames else 'no'})")


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from pero_pretraining_b200 import ops
    ops.require_device()
    check_peer_allreduce(dev, rank, world)
    check_dp_quantizer(dev, rank, world)
    check_dp_head(dev, rank, world)
    check_sharded_codebook(dev, rank, world)
    torch.cuda.synchronize()
    dist.barrier()
    log("MULTIGPU OK")
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
