"""Codebook FIT for the Feature-Quantization / Post-Quantized-AE labellers on the B200.

The reference fits its k-means codebooks on the CPU with scikit-learn (``scripts/fit_kmeans.py:20-32``:
``MiniBatchKMeans(n_clusters=k, init="k-means++", batch_size=2**14, max_iter=epochs, n_init=10).fit(vectors)``)
and pickles the estimator; ``scripts/produce_kmeans_labels.py:101`` then expects the centres as a ``.npy`` array
(the export step between the two is missing upstream — ``save_centers`` below provides it).

Here every mini-batch step runs on the device out of the kernels of the hot path:
    pero_vq_assign (tcgen05 distance GEMM + arg-min)  ->  pero_vq_ema_accumulate (deterministic segmented sum)
    ->  pero_kmeans_update (scikit-learn's count-weighted running mean, _k_means_minibatch.pyx)
The loop around it restates ``MiniBatchKMeans.fit`` of scikit-learn 1.9 (uniform batches with replacement,
low-count centre reassignment, EWA-inertia early stopping).  Randomness comes from a ``numpy`` RandomState, as in
scikit-learn, so runs are reproducible, but the individual draws differ from scikit-learn's (the inverse-CDF
sampling runs on the device); parity is pinned on the deterministic part: ``partial_fit`` from given centres on given batches
(tests/golden/kmeans_minibatch.npz, produced by scikit-learn itself).
"""
import numpy as np
import torch

from . import ops


class MiniBatchKMeans:
    """The subset of ``sklearn.cluster.MiniBatchKMeans`` the reference uses, on the B200."""

    def __init__(self, n_clusters=4096, init="k-means++", batch_size=2 ** 14, max_iter=100, n_init=1, random_state=None,
                 reassignment_ratio=0.01, max_no_improvement=10, tol=0.0, init_size=None, device=None, verbose=False):
        self.n_clusters = int(n_clusters)
        self.init = init
        self.batch_size = int(batch_size)
        self.max_iter = int(max_iter)
        self.n_init = int(n_init)
        self.reassignment_ratio = float(reassignment_ratio)
        self.max_no_improvement = max_no_improvement
        self.tol = float(tol)
        self.init_size = init_size
        self.verbose = verbose
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._rng = random_state if isinstance(random_state, np.random.RandomState) else np.random.RandomState(random_state)
        self._centers = None          # [K, D] fp32 on the device
        self._counts = None           # [K] fp32 (scikit-learn's weight_sums / _counts)
        self._codebook = None
        self.n_steps_ = 0
        self.inertia_ = None
        self._n_since_last_reassign = 0
        self.check_every = 8          # fit(): the batch inertias are read back (one host sync) every this many steps

    # ------------------------------------------------------------------------------------------ state
    @property
    def cluster_centers_(self):
        return self._centers.cpu().numpy()

    @property
    def counts_(self):
        return self._counts.cpu().numpy()

    def save_centers(self, path):
        """``np.save`` of the [K, D] float32 centres: the file ``produce_kmeans_labels.py --kmeans-path`` loads (:101)."""
        with open(path, "wb") as f:
            np.save(f, self.cluster_centers_.astype(np.float32))

    def _set_centers(self, centers):
        c = torch.as_tensor(centers, dtype=torch.float32).to(self.device).contiguous()
        if c.shape[0] != self.n_clusters:
            raise ValueError(f"init has {c.shape[0]} centres, n_clusters is {self.n_clusters}")
        self._centers = c.clone()
        self._counts = torch.zeros(self.n_clusters, dtype=torch.float32, device=self.device)
        self._codebook = ops.PreparedCodebook(self.n_clusters, c.shape[1], self.device).prepare(self._centers)

    def _as_device_rows(self, X):
        t = torch.as_tensor(X)
        if t.dtype != torch.float32:
            t = t.float()
        return t.to(self.device, non_blocking=True).contiguous()

    # ------------------------------------------------------------------------------------------ initialisation
    def _init_centroids(self, X_host_or_dev, n_samples):
        if not isinstance(self.init, str):
            self._set_centers(self.init)
            return
        init_size = self.init_size or 3 * self.batch_size
        init_size = min(max(init_size, self.n_clusters), n_samples)
        idx = self._rng.randint(0, n_samples, init_size)
        Xi = self._as_device_rows(X_host_or_dev[idx] if not torch.is_tensor(X_host_or_dev) else
                                  X_host_or_dev[torch.from_numpy(idx).to(X_host_or_dev.device)])
        if self.init == "random":
            seeds = self._rng.permutation(init_size)[:self.n_clusters]
            self._set_centers(Xi[torch.from_numpy(seeds).to(self.device)])
            return
        if self.init != "k-means++":
            raise ValueError(f"unknown init {self.init!r}")
        # Greedy k-means++ as scikit-learn runs it (_kmeans_plusplus: D^2 sampling with 2 + log(K) local trials per
        # step, keeping the candidate that lowers the potential most).  Initialisation runs once on a subsample and
        # is not on the hot path: plain tensor ops.
        K, D = self.n_clusters, Xi.shape[1]
        trials = 2 + int(np.log(K))
        centers = torch.empty(K, D, device=self.device)
        first = int(self._rng.randint(0, init_size))
        centers[0] = Xi[first]
        x2 = (Xi * Xi).sum(1)
        closest = (x2 - 2.0 * (Xi @ centers[0]) + (centers[0] * centers[0]).sum()).clamp_min_(0)
        for k in range(1, K):
            # inverse-CDF draws on the device; the uniform numbers come from the host generator (reproducible)
            u = torch.from_numpy(self._rng.random_sample(trials)).to(self.device)
            cdf = torch.cumsum(closest.double(), 0)
            picks = torch.searchsorted(cdf, cdf[-1] * u).clamp_max_(init_size - 1)
            cand = Xi[picks]                                                              # [trials, D]
            d = (x2[None, :] - 2.0 * (cand @ Xi.t()) + (cand * cand).sum(1)[:, None]).clamp_min_(0)
            d = torch.minimum(d, closest[None, :])
            best = int(torch.argmin(d.sum(1)))
            centers[k] = cand[best]
            closest = d[best]
        self._set_centers(centers)

    # ------------------------------------------------------------------------------------------ one step
    def _random_reassign(self):
        """MiniBatchKMeans._random_reassign (scikit-learn 1.9)."""
        self._n_since_last_reassign += self.batch_size
        if bool((self._counts == 0).any()) or self._n_since_last_reassign >= 10 * self.n_clusters:
            self._n_since_last_reassign = 0
            return True
        return False

    def _step(self, Xb, random_reassign):
        """_mini_batch_step: labels + inertia with the centres BEFORE the update, then the update.
        Returns the batch inertia as a 0-dim device tensor."""
        N, D = Xb.shape
        idx, _, _ = ops.vq_assign(Xb, self._codebook, N, 1, channels_first=False)
        nearest = ops.vq_gather_st(Xb, idx, self._centers, N, 1, channels_first=False)       # x + (c[idx] - x)
        inertia = ops.mse_fwd(nearest, Xb, float(N * D), 0.0)                                 # sum of squared distances
        sums = ops.vq_ema_accumulate(Xb, idx, self.n_clusters)
        ops.kmeans_update(sums, self._centers, self._counts, self._codebook)
        if random_reassign and self.reassignment_ratio > 0:
            self._reassign(Xb)
        return inertia

    def _reassign(self, Xb):
        """Low-count centres are moved onto random observations of the batch (_mini_batch_step, second half)."""
        counts = self._counts
        to_reassign = counts < self.reassignment_ratio * counts.max()
        n_batch = Xb.shape[0]
        if int(to_reassign.sum()) > 0.5 * n_batch:
            keep = torch.argsort(counts)[int(0.5 * n_batch):]
            to_reassign[keep] = False
        n = int(to_reassign.sum())
        if n:
            pick = self._rng.choice(n_batch, replace=False, size=n)
            self._centers[to_reassign] = Xb[torch.from_numpy(pick).to(self.device)]
            self._codebook.prepare(self._centers)
        if n and bool((~to_reassign).any()):
            counts[to_reassign] = counts[~to_reassign].min()

    def partial_fit(self, X):
        """One mini-batch step on the rows of X (MiniBatchKMeans.partial_fit)."""
        Xb = self._as_device_rows(X)
        if self._centers is None:
            self._init_centroids(Xb, Xb.shape[0])
        inertia = self._step(Xb, self._random_reassign())
        self.n_steps_ += 1
        self.inertia_ = float(inertia.item())
        return self

    # ------------------------------------------------------------------------------------------ fit
    def _batch(self, X, bidx, on_device):
        if on_device:
            return X[torch.from_numpy(bidx).to(X.device)].float().contiguous()
        if torch.is_tensor(X):
            return self._as_device_rows(X[torch.from_numpy(bidx)])
        return self._as_device_rows(X[bidx])

    def fit(self, X):
        """MiniBatchKMeans.fit of scikit-learn 1.9: `n_init` initialisations scored by their inertia on a validation
        subsample (the best one is kept), then max_iter passes' worth of uniformly sampled batches with the two
        early-stopping rules: EWA inertia without improvement for `max_no_improvement` steps, and (tol > 0) squared
        centre movement below tol * mean feature variance.  The batch inertias stay on the device and are read back every
        `check_every` steps, so the EWA rule may run up to check_every - 1 steps longer than scikit-learn would."""
        n_samples = len(X)
        on_device = torch.is_tensor(X) and X.is_cuda
        if not on_device and not torch.is_tensor(X):
            X = np.ascontiguousarray(X, dtype=np.float32)
        bs = min(self.batch_size, n_samples)
        self.batch_size = bs
        init_size = min(max(self.init_size or 3 * bs, self.n_clusters), n_samples)
        n_init = self.n_init if isinstance(self.init, str) else 1       # explicit centres: a single "initialisation"
        if n_init > 1:
            Xv = self._batch(X, self._rng.randint(0, n_samples, init_size), on_device)
            best = None
            for _ in range(n_init):
                self._init_centroids(X, n_samples)
                inertia = self.score_inertia(Xv)
                if best is None or inertia < best[0]:
                    best = (inertia, self._centers.clone())
            self._set_centers(best[1])
        else:
            self._init_centroids(X, n_samples)
        tol_abs = 0.0
        if self.tol > 0.0:       # scikit-learn's _tolerance: tol * mean of the per-feature variances
            Xt = X if torch.is_tensor(X) else torch.from_numpy(X)
            tol_abs = self.tol * float(Xt.float().var(dim=0, unbiased=False).mean())
        n_steps = (self.max_iter * n_samples) // bs
        ewa, ewa_min, no_improvement = None, None, 0
        pending, stop = [], False
        alpha = min(bs * 2.0 / (n_samples + 1), 1.0)
        for i in range(n_steps):
            Xb = self._batch(X, self._rng.randint(0, n_samples, bs), on_device)
            old = self._centers.clone() if tol_abs > 0.0 else None
            inertia = self._step(Xb, self._random_reassign())
            self.n_steps_ += 1
            moved = ops.mse_fwd(self._centers, old, float(old.numel()), 0.0) if old is not None else None
            pending.append((i, inertia, moved))
            if len(pending) < self.check_every and i + 1 < n_steps:
                continue
            for j, inert_t, moved_t in pending:             # one synchronisation for the whole group
                if j == 0:                  # _mini_batch_convergence ignores the first step (inertia of the init)
                    continue
                val = float(inert_t.item()) / bs
                ewa = val if ewa is None else ewa * (1 - alpha) + val * alpha
                if self.verbose:
                    print(f"Minibatch step {j + 1}/{n_steps}: mean batch inertia: {val}, ewa inertia: {ewa}")
                if moved_t is not None and float(moved_t.item()) <= tol_abs:
                    stop = True
                if ewa_min is None or ewa < ewa_min:
                    ewa_min, no_improvement = ewa, 0
                else:
                    no_improvement += 1
                if self.max_no_improvement is not None and no_improvement >= self.max_no_improvement:
                    stop = True
            pending = []
            if stop:
                break
        self.inertia_ = self.score_inertia(X)
        return self

    # ------------------------------------------------------------------------------------------ inference
    def predict(self, X, chunk=1 << 16):
        """Nearest-centre index of every row (numpy int64), in chunks."""
        out = []
        for lo in range(0, len(X), chunk):
            Xc = self._as_device_rows(X[lo:lo + chunk])
            idx, _, _ = ops.vq_assign(Xc, self._codebook, Xc.shape[0], 1, channels_first=False)
            out.append(idx.cpu())
        return torch.cat(out).numpy() if out else np.zeros(0, dtype=np.int64)

    def score_inertia(self, X, chunk=1 << 16):
        """Sum of squared distances of the rows of X to their nearest centre."""
        total = 0.0
        for lo in range(0, len(X), chunk):
            Xc = self._as_device_rows(X[lo:lo + chunk])
            n, d = Xc.shape
            idx, _, _ = ops.vq_assign(Xc, self._codebook, n, 1, channels_first=False)
            nearest = ops.vq_gather_st(Xc, idx, self._centers, n, 1, channels_first=False)
            total += float(ops.mse_fwd(nearest, Xc, float(n * d), 0.0).item())
        return total


def fit(vectors, k, batch_size=2 ** 14, epochs=100, random_state=None, device=None):
    """``scripts/fit_kmeans.py:20-32``: shuffle, fit, report the inertia, return the fitted model."""
    vectors = np.asarray(vectors, dtype=np.float32)
    rng = np.random.RandomState(random_state)
    rng.shuffle(vectors)
    kmeans = MiniBatchKMeans(n_clusters=k, init="k-means++", batch_size=batch_size, max_iter=epochs, n_init=10, random_state=rng,
                             device=device)
    kmeans.fit(vectors)
    print(f"Inertia:{kmeans.inertia_}")
    return kmeans
