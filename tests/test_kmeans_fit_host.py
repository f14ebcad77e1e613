"""Host logic of the codebook fitter (kmeans_fit.MiniBatchKMeans: scikit-learn's MiniBatchKMeans.fit loop, k-means++
initialisation, low-count reassignment, EWA early stopping) with the oracle standing in for the device kernels — the
same arrangement as the gloo tests: no GPU involved, the CUDA path itself is covered by tests/test_gpu_kmeans_fit.py."""
import types

import numpy as np
import pytest
import torch

from oracle import pero_oracle as O


class _FakeCodebook:
    def __init__(self, K, D, device):
        self.K, self.D = K, D
        self.weight = None

    def prepare(self, weight):
        self.weight = weight.detach().clone()
        return self


def _fake_ops():
    ops = types.SimpleNamespace()
    ops.PreparedCodebook = _FakeCodebook

    def vq_assign(x, cb, n_lines, frames, channels_first=False, **kw):
        return O.kmeans_assign(x, cb.weight), None, None

    def vq_gather_st(x, idx, centers, n, f, channels_first):
        return x + (centers[idx] - x)

    def mse_fwd(a, b, scale_a=1.0, scale_b=0.0):
        m = ((a - b) ** 2).mean()
        return scale_a * m + scale_b * m

    def vq_ema_accumulate(x, idx, K):
        D = x.shape[1]
        sums = torch.zeros(K, D).index_add_(0, idx, x)
        counts = torch.bincount(idx, minlength=K).float()
        return torch.cat([sums.reshape(-1), counts])

    def kmeans_update(sums_counts, centers, weight_sums, codebook=None):
        K, D = centers.shape
        sums, counts = sums_counts[:K * D].view(K, D), sums_counts[K * D:]
        hit = counts > 0
        w_new = weight_sums + counts
        centers[hit] = (centers[hit] * weight_sums[hit, None] + sums[hit]) * (1.0 / w_new[hit, None])
        weight_sums[hit] = w_new[hit]
        if codebook is not None:
            codebook.prepare(centers)

    ops.vq_assign, ops.vq_gather_st, ops.mse_fwd = vq_assign, vq_gather_st, mse_fwd
    ops.vq_ema_accumulate, ops.kmeans_update = vq_ema_accumulate, kmeans_update
    return ops


@pytest.fixture
def kmeans_cls(monkeypatch):
    from pero_pretraining_b200 import kmeans_fit
    monkeypatch.setattr(kmeans_fit, "ops", _fake_ops())
    return kmeans_fit.MiniBatchKMeans


def _blobs(K, D, n, seed, spread=5.0, noise=0.3):
    rng = np.random.RandomState(seed)
    true = rng.randn(K, D).astype(np.float32) * spread
    X = (true[rng.randint(0, K, n)] + noise * rng.randn(n, D)).astype(np.float32)
    return true, X


def test_partial_fit_host_path_matches_sklearn_golden(kmeans_cls):
    from conftest import load_golden
    g = load_golden("kmeans_minibatch")
    km = kmeans_cls(n_clusters=g["init"].shape[0], init=g["init"], reassignment_ratio=0.0, device="cpu")
    for i in range(3):
        km.partial_fit(g[f"batch{i}"])
        np.testing.assert_array_equal(km.counts_, g[f"counts{i}"])
        np.testing.assert_allclose(km.cluster_centers_, g[f"centers{i}"], rtol=1e-5, atol=1e-5)
    assert km.n_steps_ == 3 and km.inertia_ > 0


def test_fit_loop_recovers_blobs_stops_early_and_is_reproducible(kmeans_cls, tmp_path):
    true, X = _blobs(16, 8, 4000, seed=3)
    runs = []
    for _ in range(2):
        km = kmeans_cls(n_clusters=16, batch_size=256, max_iter=200, random_state=7, device="cpu").fit(X)
        runs.append(km.cluster_centers_.copy())
        assert km.n_steps_ < (200 * 4000) // 256          # the EWA-inertia rule stopped the loop early
    assert np.array_equal(runs[0], runs[1])                # same seed, same centres
    d = ((true[:, None, :] - runs[0][None]) ** 2).sum(-1)
    assert (np.sqrt(d.min(1)) < 0.5).mean() >= 0.8         # greedy k-means++ finds (nearly) every blob
    floor = 4000 * 8 * 0.3 ** 2
    assert km.inertia_ < 3.0 * floor
    p = tmp_path / "centers.npy"
    km.save_centers(p)
    c = np.load(p)
    assert c.dtype == np.float32 and c.shape == (16, 8)
    assert np.array_equal(km.predict(X[:100]), O.kmeans_assign(torch.from_numpy(X[:100]), torch.from_numpy(c)).numpy())


def test_low_count_centres_are_reassigned(kmeans_cls):
    """A centre far away from all data never wins a frame; the reassignment step (scikit-learn's _mini_batch_step,
    second half) must move it onto an observation."""
    true, X = _blobs(4, 6, 2000, seed=5)
    init = np.concatenate([true[:3] + 0.1, np.full((1, 6), 1e3, np.float32)]).astype(np.float32)
    km = kmeans_cls(n_clusters=4, init=init, batch_size=500, max_iter=5, random_state=0, device="cpu").fit(X)
    assert np.abs(km.cluster_centers_).max() < 100.0       # the stray centre was pulled back onto the data
    assert (km.counts_ > 0).all()
    frozen = kmeans_cls(n_clusters=4, init=init, batch_size=500, max_iter=5, random_state=0, reassignment_ratio=0.0,
                        device="cpu").fit(X)
    assert np.abs(frozen.cluster_centers_).max() > 100.0   # without reassignment it stays where it was


def test_init_validation(kmeans_cls):
    with pytest.raises(ValueError):
        kmeans_cls(n_clusters=5, init=np.zeros((4, 3), np.float32), device="cpu").partial_fit(np.zeros((10, 3), np.float32))
    with pytest.raises(ValueError):
        kmeans_cls(n_clusters=4, init="bogus", device="cpu").fit(np.random.rand(50, 3).astype(np.float32))
