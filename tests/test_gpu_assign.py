"""GPU parity of the tcgen05 GEMM core and the nearest-codeword assignment against the oracle
(oracle/pero_oracle.py) and the golden fixtures.  All calls go through the C ABI (ctypes).

Near-tie rule (north_star): a CUDA index may differ from the reference index only on frames whose fp64
relative top-2 distance gap (d2 - d1) / d1 is below EPS_TIE.  The GEMM uses bf16 operands with fp32
accumulation; measured flips on N(0,1) data all have gap < 6e-4 (median gap ~1e-2)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import pero_oracle as O

pytestmark = pytest.mark.gpu

EPS_TIE = 2e-3


def _ops():
    from pero_pretraining_b200 import ops
    return ops


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
@pytest.mark.parametrize("shape", [(128, 256, 64), (300, 700, 256), (1000, 1000, 512), (129, 257, 128)])
def test_gemm_core_matches_matmul(cuda_dev, variant, shape):
    ops = _ops()
    ra, rb, kd = shape
    if variant == 2 and kd > 256:
        pytest.skip("plain-store epilogue (37 KB staging) + 128 KB resident A + 32 KB stages exceeds one CTA's shared memory")
    g = torch.Generator(device="cpu").manual_seed(ra * 7 + rb)
    a = torch.randn(ra, kd, generator=g).to(cuda_dev).bfloat16()
    b = torch.randn(rb, kd, generator=g).to(cuda_dev).bfloat16()
    got = ops.debug_gemm_tn(a, b, variant=variant)[0]
    ref = a.double() @ b.double().t()
    # bf16 products are exact in fp32; only the fp32 accumulation order differs
    assert (got.double() - ref).abs().max().item() <= 1e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("pairs", [0, 1])
@pytest.mark.parametrize("shape", [(128, 256, 64), (304, 520, 200), (1000, 512, 1245), (8192, 512, 150)])
def test_gemm_core_mn_major_operands(cuda_dev, pairs, shape):
    """C = A^T B with A stored [k, rows_a] and B [k, rows_b] (MN-major UMMA descriptors fed by {64 x 64} TMA boxes):
    how d_W = dlogits^T h is computed without a transposed copy of either operand."""
    ops = _ops()
    ra, rb, kd = shape
    g = torch.Generator(device="cpu").manual_seed(ra + rb + kd)
    a = torch.randn(kd, ra, generator=g).to(cuda_dev).bfloat16()
    b = torch.randn(kd, rb, generator=g).to(cuda_dev).bfloat16()
    out = torch.zeros(1, ra, rb, dtype=torch.float32, device=cuda_dev)
    from pero_pretraining_b200 import _lib
    _lib.check(_lib.lib().pero_gemm_tn_bf16(a.data_ptr(), ra, b.data_ptr(), rb, kd, 32 + pairs, 1, out.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream), "pero_gemm_tn_bf16")
    ref = a.double().t() @ b.double()
    assert (out[0].double() - ref).abs().max().item() <= 1e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("variant", [0, 1])
def test_gemm_core_streamed_long_k_and_splits(cuda_dev, variant):
    ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(5)
    a = torch.randn(300, 1280, generator=g).to(cuda_dev).bfloat16()
    b = torch.randn(520, 1280, generator=g).to(cuda_dev).bfloat16()
    ref = a.double() @ b.double().t()
    for splits in (1, 3, 20):
        got = ops.debug_gemm_tn(a, b, variant=variant, splits=splits).sum(0)
        assert (got.double() - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()


def _assign_vs_oracle(ops, x_rows, w, channels_first_x=None, n_lines=None, frames=None):
    K, D = w.shape
    cb = ops.PreparedCodebook(K, D, w.device).prepare(w)
    if channels_first_x is not None:
        idx, dmin, rows = ops.vq_assign(channels_first_x, cb, n_lines, frames, True, want_dmin=True, want_rows=True)
        assert torch.equal(rows, x_rows)                      # the fp32 row copy is a pure transpose
    else:
        idx, dmin, _ = ops.vq_assign(x_rows, cb, x_rows.shape[0], 1, False, want_dmin=True)
    torch.cuda.synchronize()
    ref_idx, ref_dmin, gap = O.assign_fp64(x_rows.cpu().numpy(), w.cpu().numpy())
    got = idx.cpu().numpy()
    differs = got != ref_idx
    assert (gap[differs] < EPS_TIE).all(), f"{differs.sum()} flips, worst gap {gap[differs].max():.3e}"
    assert differs.mean() < 0.01
    # dmin is the squared distance minus |x|^2, evaluated on bf16-rounded operands
    xn = (x_rows.double() ** 2).sum(1).cpu().numpy()
    np.testing.assert_allclose(dmin.cpu().numpy()[~differs] + xn[~differs], ref_dmin[~differs],
                               rtol=0, atol=2e-2 * np.abs(ref_dmin).max())
    return got, ref_idx


@pytest.mark.parametrize("N,K,D", [(1, 1, 8), (5, 3, 8), (700, 1000, 200), (1024, 4096, 512), (257, 300, 768),
                                   (2000, 513, 64), (128, 256, 1000)])
def test_assign_rows_vs_fp64_oracle(cuda_dev, N, K, D):
    ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(N + K + D)
    x = torch.randn(N, D, generator=g).to(cuda_dev)
    w = torch.randn(K, D, generator=g).to(cuda_dev)
    _assign_vs_oracle(ops, x, w)


def test_assign_channels_first_mixture_c1(cuda_dev):
    """8 lines x 128 frames, 4096 x 512 codebook, frames = codeword + 0.5 N(0,1)."""
    ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(1235)
    C = torch.randn(4096, 512, generator=g)
    j = torch.randint(0, 4096, (1024,), generator=g)
    rows = C[j] + 0.5 * torch.randn(1024, 512, generator=g)
    x = rows.view(8, 1, 128, 512).permute(0, 3, 1, 2).contiguous()
    got, ref = _assign_vs_oracle(ops, rows.to(cuda_dev), C.to(cuda_dev), x.view(8, 512, 128).to(cuda_dev), 8, 128)
    assert (got == j.numpy()).mean() > 0.99      # realistic gaps: the generating codeword wins


def test_assign_golden_kmeans_fixture(cuda_dev):
    """Labels of scripts/produce_kmeans_labels.py:72-80 as produced by the reference's own torch ops."""
    from pero_pretraining_b200 import KMeansLabeller
    g = load_golden("kmeans_assign")
    lab = KMeansLabeller(torch.from_numpy(g["centers"]).to(cuda_dev)).assign_features(torch.from_numpy(g["features"]).to(cuda_dev))
    got = lab.cpu().numpy()
    f = torch.from_numpy(g["features"]).squeeze(2).permute(0, 2, 1).reshape(-1, g["centers"].shape[1])
    _, _, gap = O.assign_fp64(f.numpy(), g["centers"])
    differs = (got != g["labels"]).reshape(-1)
    assert (gap[differs] < EPS_TIE).all()
    assert got.shape == g["labels"].shape and got.dtype == np.int64


def test_assign_ties_pick_lowest_index(cuda_dev):
    """torch.argmin returns the first minimal index (autoencoders.py:217): duplicate codewords give exactly
    equal distances in every K-tile / CTA split, so the smallest duplicate must win."""
    ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(3)
    base = torch.randn(600, 64, generator=g)
    w = torch.cat([base, base, base], 0).to(cuda_dev)        # K = 1800: duplicates 600 and 1200 apart
    x = (base[torch.randint(0, 600, (900,), generator=g)] + 0.01 * torch.randn(900, 64, generator=g)).to(cuda_dev)
    cb = ops.PreparedCodebook(1800, 64, cuda_dev).prepare(w)
    idx, _, _ = ops.vq_assign(x, cb, 900, 1, False)
    assert int(idx.max()) < 600
    ref = O.vq_assign_fp32(x.cpu(), w.cpu())
    assert (idx.cpu() == ref).float().mean() > 0.99


def test_assign_fixed_point_and_full_size_properties(cuda_dev):
    """Size-independent properties at BASELINE config sizes (the fp64 oracle is too slow there):
    c2 (8192 frames, 8192 x 256): every codeword is its own nearest codeword (idempotence);
    c4 (65536 frames, 16384 x 512): frames generated around codewords recover the generating index, and the
    returned dmin equals the distance recomputed in fp64 for the returned index."""
    ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(1236)
    C = torch.randn(8192, 256, generator=g).to(cuda_dev)
    cb = ops.PreparedCodebook(8192, 256, cuda_dev).prepare(C)
    idx, _, _ = ops.vq_assign(C, cb, 8192, 1, False)
    assert torch.equal(idx, torch.arange(8192, device=cuda_dev))
    # channels-first entry gives the same answer as the row entry
    xcf = C.view(64, 128, 256).permute(0, 2, 1).contiguous()
    idx2, _, _ = ops.vq_assign(xcf, cb, 64, 128, True)
    assert torch.equal(idx2, idx)

    C4 = torch.randn(16384, 512, generator=g).to(cuda_dev)
    j = torch.randint(0, 16384, (65536,), generator=g).to(cuda_dev)
    X = C4[j] + 0.5 * torch.randn(65536, 512, generator=g).to(cuda_dev)
    cb4 = ops.PreparedCodebook(16384, 512, cuda_dev).prepare(C4)
    idx4, dmin4, _ = ops.vq_assign(X, cb4, 65536, 1, False, want_dmin=True)
    assert (idx4 == j).float().mean().item() > 0.999
    d = ((X.double() - C4[idx4].double()) ** 2).sum(1) - (X.double() ** 2).sum(1)
    assert (dmin4.double() - d).abs().max().item() < 2e-2 * d.abs().max().item()


def _near_tie_check(idx, x_rows, w, sample=None, seed=0):
    """Near-tie rule against the fp64 brute force run ON THE DEVICE (oracle.assign_fp64_torch) for a random subset
    of the frames (the assignment is row-independent, so a subset of rows is an exact check of those rows).
    Returns (flip rate, worst gap among the flips)."""
    N = x_rows.shape[0]
    if sample is not None and sample < N:
        g = torch.Generator(device="cpu").manual_seed(seed)
        sel = torch.randperm(N, generator=g)[:sample].to(x_rows.device)
    else:
        sel = torch.arange(N, device=x_rows.device)
    ref_idx, _, gap = O.assign_fp64_torch(x_rows[sel], w)
    differs = idx[sel] != ref_idx
    worst = float(gap[differs].max()) if bool(differs.any()) else 0.0
    assert worst < EPS_TIE, f"{int(differs.sum())} flips, worst gap {worst:.3e}"
    rate = float(differs.float().mean())
    assert rate <= 0.01, f"flip rate {rate:.4f}"
    return rate, worst


def test_device_fp64_oracle_matches_numpy_oracle(cuda_dev):
    g = torch.Generator(device="cpu").manual_seed(11)
    x, w = torch.randn(700, 96, generator=g), torch.randn(333, 96, generator=g)
    i_np, d_np, gap_np = O.assign_fp64(x.numpy(), w.numpy())
    i_t, d_t, gap_t = O.assign_fp64_torch(x.to(cuda_dev), w.to(cuda_dev), chunk=128)
    assert np.array_equal(i_t.cpu().numpy(), i_np)
    np.testing.assert_allclose(d_t.cpu().numpy(), d_np, rtol=1e-10)
    np.testing.assert_allclose(gap_t.cpu().numpy(), gap_np, rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("kind", ["normal", "mixture"])
def test_assign_near_tie_rule_config2_full_size(cuda_dev, kind):
    """BASELINE configs[1] (8192 frames, 8192 x 256 codebook) against the fp64 oracle on every frame: pure N(0,1)
    frames are the worst case SURVEY 8d names (median relative top-2 gap ~1e-2), the mixture is the bench's input."""
    ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(1236)
    C = torch.randn(8192, 256, generator=g).to(cuda_dev)
    if kind == "normal":
        X = torch.randn(8192, 256, generator=g).to(cuda_dev)
    else:
        j = torch.randint(0, 8192, (8192,), generator=g).to(cuda_dev)
        X = C[j] + 0.5 * torch.randn(8192, 256, generator=g).to(cuda_dev)
    cb = ops.PreparedCodebook(8192, 256, cuda_dev).prepare(C)
    idx, _, _ = ops.vq_assign(X, cb, 8192, 1, False)
    xcf = X.view(64, 128, 256).permute(0, 2, 1).contiguous()
    idx_cf, _, _ = ops.vq_assign(xcf, cb, 64, 128, True)
    assert torch.equal(idx, idx_cf)
    rate, worst = _near_tie_check(idx, X, C)
    print(f"config 2 ({kind}): flip rate {rate:.5f}, worst flipped gap {worst:.2e}")


def test_assign_near_tie_rule_config4_full_size(cuda_dev):
    """BASELINE configs[3] (65536 frames, 16384 x 512): all frames assigned, 8192 random frames checked in fp64;
    both the mixture and pure N(0,1) frames."""
    ops = _ops()
    g = torch.Generator(device=cuda_dev).manual_seed(1238)
    K, D, N = 16384, 512, 65536
    C = torch.randn(K, D, device=cuda_dev, generator=g)
    cb = ops.PreparedCodebook(K, D, cuda_dev).prepare(C)
    j = torch.randint(0, K, (N,), device=cuda_dev, generator=g)
    for kind in ("mixture", "normal"):
        X = torch.randn(N, D, device=cuda_dev, generator=g)
        if kind == "mixture":
            X = C[j] + 0.5 * X
        idx, _, _ = ops.vq_assign(X, cb, N, 1, False)
        rate, worst = _near_tie_check(idx, X, C, sample=8192, seed=4)
        print(f"config 4 ({kind}): flip rate {rate:.5f}, worst flipped gap {worst:.2e}")


def test_assign_near_tie_rule_config5_subsample(cuda_dev):
    """BASELINE configs[4] codebook (65536 x 512): 65536 frames assigned against the whole codebook and against 8
    shards merged by MIN (bit-identical), 4096 random frames checked in fp64."""
    ops = _ops()
    g = torch.Generator(device=cuda_dev).manual_seed(1239)
    K, D, N, world = 65536, 512, 65536, 8
    C = torch.randn(K, D, device=cuda_dev, generator=g)
    X = torch.randn(N, D, device=cuda_dev, generator=g)
    full = ops.PreparedCodebook(K, D, cuda_dev).prepare(C)
    idx, _, _ = ops.vq_assign(X, full, N, 1, False)
    packed = ops.vq_packed_init(N, cuda_dev)
    for r in range(world):
        lo = r * (K // world)
        shard = ops.PreparedCodebook(K // world, D, cuda_dev).prepare(C[lo:lo + K // world].contiguous())
        ops.vq_assign(X, shard, N, 1, False, index_offset=lo, packed=packed)
    idx_sh, _ = ops.vq_unpack(packed)
    assert torch.equal(idx, idx_sh)
    rate, worst = _near_tie_check(idx, X, C, sample=4096, seed=5)
    print(f"config 5 (N(0,1) frames): flip rate {rate:.5f}, worst flipped gap {worst:.2e}")


def test_config5_codebook_sharded_stress_full_size(cuda_dev):
    """BASELINE config 5 at full size on one GPU: 2**20 frames, 65536 x 512 codebook walked as 8 shards of 8192
    codewords whose packed winners are min-merged (what the 8-rank MIN exchange does).  Frames generated around
    codewords recover the generating index, and the merge is bit-identical to one call on the whole codebook."""
    import time
    ops = _ops()
    g = torch.Generator(device=cuda_dev).manual_seed(5)
    K, D, N, world = 65536, 512, 1 << 20, 8
    C = torch.randn(K, D, device=cuda_dev, generator=g)
    j = torch.randint(0, K, (N,), device=cuda_dev, generator=g)
    X = C[j]
    X += 0.5 * torch.randn(N, D, device=cuda_dev, generator=g)
    packed = ops.vq_packed_init(N, cuda_dev)
    shards = [ops.PreparedCodebook(K // world, D, cuda_dev).prepare(C[r * (K // world):(r + 1) * (K // world)].contiguous())
              for r in range(world)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for r in range(world):
        ops.vq_assign(X, shards[r], N, 1, False, index_offset=r * (K // world), packed=packed)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    idx, _ = ops.vq_unpack(packed)
    assert (idx == j).float().mean().item() > 0.999
    full = ops.PreparedCodebook(K, D, cuda_dev).prepare(C)
    packed_full = ops.vq_packed_init(N, cuda_dev)
    ops.vq_assign(X, full, N, 1, False, packed=packed_full)
    assert torch.equal(packed_full, packed)
    print(f"config 5 on one GPU: {dt * 1e3:.1f} ms for {2.0 * N * K * D / 1e12:.1f} TFLOP "
          f"({2.0 * N * K * D / dt / 1e12:.0f} TFLOP/s incl. frame preparation)")


def test_codebook_sharded_merge_equals_full_assign(cuda_dev):
    """SURVEY §8e codebook-sharded mode on one GPU: g shards packed into the same int64 buffer by atomicMin
    (what the MIN all-reduce does across ranks) == assignment against the whole codebook."""
    ops = _ops()
    from pero_pretraining_b200.sharding import shard_bounds
    g = torch.Generator(device="cpu").manual_seed(9)
    K, D, N = 3000, 128, 2048
    w = torch.randn(K, D, generator=g).to(cuda_dev)
    x = torch.randn(N, D, generator=g).to(cuda_dev)
    full = ops.PreparedCodebook(K, D, cuda_dev).prepare(w)
    idx_full, dmin_full, _ = ops.vq_assign(x, full, N, 1, False, want_dmin=True)
    for world in (2, 8):
        merged = None
        for r in range(world):
            lo, hi = shard_bounds(K, world, r)
            shard = ops.PreparedCodebook(hi - lo, D, cuda_dev).prepare(w[lo:hi].contiguous())
            packed = ops.vq_packed_init(N, cuda_dev)
            ops.vq_assign(x, shard, N, 1, False, index_offset=lo, packed=packed)
            merged = packed if merged is None else torch.minimum(merged, packed)     # int64 MIN, as the all-reduce
        idx, dmin = ops.vq_unpack(merged, want_dmin=True)
        assert torch.equal(idx, idx_full)
        assert torch.equal(dmin, dmin_full)


def test_assign_rejects_bad_arguments(cuda_dev):
    from pero_pretraining_b200 import PeroError, _lib
    ops = _ops()
    w = torch.randn(16, 8, device=cuda_dev)
    cb = ops.PreparedCodebook(16, 8, cuda_dev).prepare(w)
    with pytest.raises(TypeError):
        ops.vq_assign(torch.randn(4, 8), cb, 4, 1, False)                 # CPU tensor: no fallback
    L = _lib.lib()
    assert L.pero_vq_assign(None, 4, 1, 0, 16, 8, cb.blob.data_ptr(), 0, None, None, None, None, None, 0, None) == -5
    x = torch.randn(4, 8, device=cuda_dev)
    ws = torch.empty(256, dtype=torch.uint8, device=cuda_dev)
    rc = L.pero_vq_assign(x.data_ptr(), 4, 1, 0, 16, 8, cb.blob.data_ptr(), 0, None, None, None, None, ws.data_ptr(), 16, None)
    assert rc == -3 and b"workspace" in L.pero_strerror(rc)
    with pytest.raises(PeroError):
        _lib.check(rc, "assign")
    # empty input is a no-op, like the reference on a zero-frame batch
    idx, _, _ = ops.vq_assign(torch.empty(0, 8, device=cuda_dev), cb, 0, 1, False)
    assert idx.numel() == 0
