"""Peer-memory exchange kernels (include/pero_b200.h "peer-memory collectives").

Single-GPU: the protocol (flag handshakes, slice ownership, in-place reduce + broadcast, replayability) is run
with the `world` ranks played by blockIdx.y of ONE cooperative launch over `world` buffers on one device.
Multi-GPU (skipped unless >= 2 devices): tests/multigpu_worker.py under torchrun — the real NVLink / NVSwitch
path, data-parallel and codebook-sharded equivalence with the single-process result."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = 16384


def _buffers(world, payload_bytes, dev):
    bufs = [torch.zeros(HEADER + payload_bytes, dtype=torch.uint8, device=dev) for _ in range(world)]
    return bufs


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("n", [4, 1000, 8192 * 256 + 8192])
def test_emulated_sum_is_rank_ordered_and_replicated(cuda_dev, world, n):
    from pero_pretraining_b200.peer import emulate_all_reduce
    g = torch.Generator(device="cpu").manual_seed(world * 1000 + n % 997)
    vals = [torch.randn(n, generator=g) * (10.0 ** (r % 3)) for r in range(world)]
    bufs = _buffers(world, 4 * n + 256, cuda_dev)
    off = HEADER + 256
    views = [b[off:off + 4 * n].view(torch.float32) for b in bufs]
    want = vals[0].clone()
    for r in range(1, world):
        want = want + vals[r]                           # fp32, rank order 0..g-1: what the owner computes
    for rep in range(3):                                # epochs only grow: the same buffers are reusable
        for v, src in zip(views, vals):
            v.copy_(src)
        emulate_all_reduce(bufs, "sum", off, n, n_blocks=1 + rep * 3)
        for r in range(world):
            assert torch.equal(views[r].cpu(), want), f"rank {r} rep {rep}"
        for b in bufs:                                      # every block that ran advanced its epoch by 2 per launch
            assert int(b[8192:8196].view(torch.int32)) == 2 * (rep + 1)


@pytest.mark.parametrize("world", [2, 5, 8])
def test_emulated_min_i64(cuda_dev, world):
    from pero_pretraining_b200.peer import emulate_all_reduce
    n = 6002
    g = torch.Generator(device="cpu").manual_seed(world)
    vals = [torch.randint(-2 ** 62, 2 ** 62, (n,), generator=g, dtype=torch.int64) for _ in range(world)]
    vals[0][:5] = torch.iinfo(torch.int64).max          # "empty" winners lose against anything
    bufs = _buffers(world, 8 * n, cuda_dev)
    views = [b[HEADER:HEADER + 8 * n].view(torch.int64) for b in bufs]
    for v, src in zip(views, vals):
        v.copy_(src)
    emulate_all_reduce(bufs, "min", HEADER, n, n_blocks=8)
    want = torch.stack(vals).min(0).values
    for r in range(world):
        assert torch.equal(views[r].cpu(), want)


def test_peer_argument_errors(cuda_dev):
    from pero_pretraining_b200 import _lib
    L = _lib.lib()
    bufs = _buffers(2, 4096, cuda_dev)
    ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=cuda_dev)
    s = torch.cuda.current_stream().cuda_stream
    assert L.pero_peer_allreduce_sum_f32(None, None, 0, 2, HEADER, 8, 4, s) == -5
    assert L.pero_peer_allreduce_sum_f32(ptrs.data_ptr(), None, 2, 2, HEADER, 8, 4, s) == -1       # rank out of range
    assert L.pero_peer_allreduce_sum_f32(ptrs.data_ptr(), None, 0, 2, HEADER - 16, 8, 4, s) == -2   # inside the header
    assert L.pero_peer_allreduce_sum_f32(ptrs.data_ptr(), None, 0, 2, HEADER, 6, 4, s) == -2        # not a multiple of 4
    assert L.pero_peer_allreduce_sum_f32(ptrs.data_ptr(), None, 0, 2, HEADER, 8, 65, s) == -1       # too many CTAs
    assert L.pero_peer_allreduce_min_i64(ptrs.data_ptr(), None, 0, 2, HEADER, 3, 4, s) == -2
    assert L.pero_peer_allreduce_sum_f32(ptrs.data_ptr(), None, 0, 1, HEADER, 8, 4, s) == 0         # world 1: nothing to do
    torch.cuda.synchronize()


def test_multi_gpu_equivalence(cuda_dev):
    """2 ranks on 2 GPUs: peer all-reduce == reference sums, data-parallel VQ/CE == single process on the
    concatenated batch, codebook-sharded assign == full assign."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run under `gpurun --gpus 2`)")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tests", "multigpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    sys.stdout.write(res.stdout[-6000:])
    sys.stderr.write(res.stderr[-6000:])
    assert res.returncode == 0
    assert "MULTIGPU OK" in res.stdout
