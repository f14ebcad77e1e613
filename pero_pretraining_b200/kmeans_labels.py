"""Feature-Quantization / Post-Quantized-AE label assignment: the math core of
``pero_pretraining/scripts/produce_kmeans_labels.py`` (:34, :72-80) -- nearest k-means centre per frame --
on the same tcgen05 distance kernel as the VQ-VAE quantizer.  cdist is a Euclidean (not squared) distance;
the arg-min is the same.
"""
import torch

from . import ops


class KMeansLabeller:
    """Holds the prepared centres [K, D] (what the script loads with np.load, :101-102)."""

    def __init__(self, centers):
        if not centers.is_cuda:
            raise ops._lib.PeroError("KMeansLabeller runs on a CUDA (B200) device only; there is no CPU path")
        self.centers = centers.detach().float().contiguous()
        K, D = self.centers.shape
        self.codebook = ops.PreparedCodebook(K, D, self.centers.device).prepare(self.centers)

    def assign_rows(self, features_linear, want_dmin=False):
        """features_linear [N, D] -> labels int64 [N]  (cdist + argmin(dim=1), :74-76)."""
        f = features_linear.detach().float().contiguous()
        idx, dmin, _ = ops.vq_assign(f, self.codebook, f.shape[0], 1, channels_first=False, want_dmin=want_dmin)
        return (idx, dmin) if want_dmin else idx

    def assign_features(self, features):
        """features [B, D, T] or [B, D, 1, T] as the encoder emits them -> labels int64 [B, T] (:54-57, :72-79).
        The channels-first layout is consumed directly; no permute/reshape copy is made."""
        if features.dim() == 4:
            features = features.squeeze(2)
        B, D, T = features.shape
        f = features.detach().float().contiguous()
        idx, _, _ = ops.vq_assign(f, self.codebook, B, T, channels_first=True)
        return idx.view(B, T)


def kmeans_assign(features_linear, centers):
    """One-shot functional form: labels = argmin_k ||f - c_k||."""
    return KMeansLabeller(centers).assign_rows(features_linear)
