"""Codebook fit on the device and the label file format / production loop (§8f-3)."""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import pero_oracle as O

pytestmark = pytest.mark.gpu


def test_partial_fit_matches_sklearn_golden(cuda_dev):
    """Same centres and counts as scikit-learn's own MiniBatchKMeans.partial_fit on the committed batches."""
    from pero_pretraining_b200 import MiniBatchKMeans
    g = load_golden("kmeans_minibatch")
    km = MiniBatchKMeans(n_clusters=g["init"].shape[0], init=g["init"], reassignment_ratio=0.0, device=cuda_dev)
    for i in range(3):
        km.partial_fit(g[f"batch{i}"])
        np.testing.assert_array_equal(km.counts_, g[f"counts{i}"])
        np.testing.assert_allclose(km.cluster_centers_, g[f"centers{i}"], rtol=5e-6, atol=5e-6)
    ref_labels, ref_inertia, _, _ = O.minibatch_kmeans_step(g["batch2"], g["centers2"], g["counts2"])
    assert np.array_equal(km.predict(g["batch2"]), ref_labels)
    assert abs(km.score_inertia(g["batch2"]) - ref_inertia) <= 1e-4 * ref_inertia


def test_partial_fit_is_deterministic_and_matches_oracle_on_a_large_batch(cuda_dev):
    from pero_pretraining_b200 import MiniBatchKMeans
    rng = np.random.RandomState(3)
    K, D, n = 512, 256, 16384                      # a reference-sized batch (2**14) through the CUB sort path
    true = rng.randn(K, D).astype(np.float32) * 3
    init = (true + 0.2 * rng.randn(K, D)).astype(np.float32)
    X = (true[rng.randint(0, K, n)] + 0.5 * rng.randn(n, D)).astype(np.float32)
    runs = []
    for _ in range(2):
        km = MiniBatchKMeans(n_clusters=K, init=init, reassignment_ratio=0.0, device=cuda_dev)
        km.partial_fit(X).partial_fit(X[::-1].copy())
        runs.append(km.cluster_centers_.copy())
    assert np.array_equal(runs[0], runs[1])
    _, _, c, w = O.minibatch_kmeans_step(X, init, np.zeros(K, np.float32))
    _, _, c, w = O.minibatch_kmeans_step(X[::-1].copy(), c, w)
    np.testing.assert_allclose(runs[0], c, rtol=2e-5, atol=2e-5)
    np.testing.assert_array_equal(km.counts_, w)


def test_fit_honours_n_init_and_tol(cuda_dev):
    """n_init initialisations scored on a validation subsample (scripts/fit_kmeans.py:21 fits with n_init=10) and
    scikit-learn's tol rule (squared centre movement below tol * mean feature variance stops the loop)."""
    from pero_pretraining_b200 import MiniBatchKMeans
    rng = np.random.RandomState(11)
    K, D, n = 32, 16, 8000
    true = rng.randn(K, D).astype(np.float32) * 4
    X = (true[rng.randint(0, K, n)] + 0.3 * rng.randn(n, D)).astype(np.float32)
    one = MiniBatchKMeans(n_clusters=K, batch_size=1024, max_iter=10, n_init=1, random_state=5, device=cuda_dev).fit(X)
    ten = MiniBatchKMeans(n_clusters=K, batch_size=1024, max_iter=10, n_init=10, random_state=5, device=cuda_dev).fit(X)
    assert ten.inertia_ <= 1.3 * one.inertia_
    loose = MiniBatchKMeans(n_clusters=K, batch_size=1024, max_iter=50, n_init=1, tol=0.5, max_no_improvement=None,
                            random_state=5, device=cuda_dev).fit(X)
    full = MiniBatchKMeans(n_clusters=K, batch_size=1024, max_iter=50, n_init=1, tol=0.0, max_no_improvement=None,
                           random_state=5, device=cuda_dev).fit(X)
    assert loose.n_steps_ < full.n_steps_ == (50 * n) // 1024


def test_fit_recovers_blobs_and_exports_centres(cuda_dev, tmp_path):
    from pero_pretraining_b200 import KMeansLabeller, MiniBatchKMeans
    from sklearn.cluster import MiniBatchKMeans as SkMBK
    rng = np.random.RandomState(7)
    K, D, n = 64, 32, 20000
    true = rng.randn(K, D).astype(np.float32) * 5
    X = (true[rng.randint(0, K, n)] + 0.3 * rng.randn(n, D)).astype(np.float32)
    km = MiniBatchKMeans(n_clusters=K, batch_size=2048, max_iter=20, random_state=0, device=cuda_dev).fit(X)
    sk = SkMBK(n_clusters=K, batch_size=2048, max_iter=20, n_init=1, random_state=0).fit(X)
    assert km.inertia_ <= 1.5 * sk.inertia_                     # same quality class as the CPU fitter on the same data
    d = ((true[:, None, :] - km.cluster_centers_[None]) ** 2).sum(-1)
    assert (np.sqrt(d.min(1)) < 0.5).mean() > 0.8               # most true blobs have a fitted centre on top of them
    path = os.path.join(tmp_path, "centers.npy")
    km.save_centers(path)
    centers = np.load(path)                                      # what produce_kmeans_labels.py:101 does
    assert centers.dtype == np.float32 and centers.shape == (K, D)
    lab = KMeansLabeller(torch.from_numpy(centers).to(cuda_dev)).assign_rows(torch.from_numpy(X[:4096]).to(cuda_dev))
    assert np.array_equal(lab.cpu().numpy(), km.predict(X[:4096]))


def test_label_file_format_and_production_loop(cuda_dev, tmp_path):
    from pero_pretraining_b200 import load_labels, produce_kmeans_labels, save_labels
    rng = np.random.RandomState(11)
    K, D, T = 96, 48, 40
    centers = rng.randn(K, D).astype(np.float32) * 3
    batches, want = [], {}
    for b in range(7):                                           # more batches than the read-back ring is deep
        B = 5 if b != 6 else 3                                   # ragged last batch
        idx = rng.randint(0, K, (B, T))
        feats = centers[idx] + 0.2 * rng.randn(B, T, D).astype(np.float32)            # [B, T, D]
        masks = np.ones((B, T), dtype=np.uint8)
        for r in range(B):
            masks[r, rng.randint(T // 2, T + 1):] = 0                                  # right padding of each line
        ids = [f"line_{b}_{r}.jpg" for r in range(B)]
        layout = torch.from_numpy(feats).permute(0, 2, 1).contiguous()                 # [B, D, T] like the encoder output
        if b % 2:
            layout = layout.unsqueeze(2)                                               # [B, D, 1, T]
        batches.append({"features": layout.to(cuda_dev), "ids": ids, "image_masks": masks})
        ref = O.kmeans_assign(torch.from_numpy(feats).reshape(-1, D), torch.from_numpy(centers)).view(B, T).numpy()
        for r in range(B):
            want[ids[r]] = ref[r][masks[r] == 1].tolist()
    path = os.path.join(tmp_path, "labels.txt")
    assert produce_kmeans_labels(iter(batches), centers, path) == len(want)
    text = open(path).read().splitlines()
    assert text[0] == f"line_0_0.jpg {' '.join(str(v) for v in want['line_0_0.jpg'])}"  # the exact wire format
    assert [ln.split()[0] for ln in text] == list(want.keys())                          # input order is kept
    assert load_labels(path) == want
    path2 = os.path.join(tmp_path, "labels2.txt")
    save_labels(want, path2)                                                            # scripts/common.py:51-54
    assert open(path2).read() == open(path).read()
