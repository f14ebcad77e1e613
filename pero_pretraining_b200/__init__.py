"""pero_pretraining_b200 — B200-native (sm_100a) quantize-and-predict path of DCGM/pero-pretraining.

Host side of the C ABI in ``include/pero_b200.h`` (``libpero_b200.so``): drop-in modules with the reference's
names and signatures.  There is no CPU or PyTorch fallback: importing works anywhere, but every operation
raises unless the shared library is built and a B200 is the current device.

    from pero_pretraining_b200 import VectorQuantizer, VQVAE                     # models/autoencoders.py
    from pero_pretraining_b200 import LinearHead, MaskedCrossEntropyLoss, MaskedTransformerEncoder
    from pero_pretraining_b200 import KMeansLabeller, kmeans_assign              # scripts/produce_kmeans_labels.py
    from pero_pretraining_b200 import MiniBatchKMeans                            # scripts/fit_kmeans.py (GPU fit)
    from pero_pretraining_b200 import produce_kmeans_labels, save_labels         # label wire format + production loop
    pero_pretraining_b200.install()   # swap the classes into an importable `pero_pretraining` package
"""
from ._lib import PeroError, build, lib  # noqa: F401
from .autoencoders import VQVAE, VectorQuantizer  # noqa: F401
from .kmeans_fit import MiniBatchKMeans  # noqa: F401
from .kmeans_labels import KMeansLabeller, kmeans_assign  # noqa: F401
from .labels_io import LabelWriter, compute_labels, load_labels, produce_kmeans_labels, save_labels  # noqa: F401
from .masked_pretraining import (LinearHead, MaskedCrossEntropyLoss, MaskedTransformerEncoder, PixelMasker,  # noqa: F401
                                 create_mask, masked_rows, update_errors)
from .sharding import ShardedCodebook, merge_packed, shard_bounds  # noqa: F401

__version__ = "0.1.0"


def install():
    """Monkey-patch the reference package (if importable) so that its trainers / scripts pick up the B200
    modules: ``pero_pretraining.models.autoencoders.{VectorQuantizer,VQVAE}`` and
    ``pero_pretraining.masked_pretraining.model.{LinearHead,MaskedCrossEntropyLoss,MaskedTransformerEncoder}``.
    Returns the list of patched attribute paths."""
    import importlib

    patched = []
    targets = {
        "pero_pretraining.models.autoencoders": {"VectorQuantizer": VectorQuantizer, "VQVAE": VQVAE},
        "pero_pretraining.masked_pretraining.model": {
            "LinearHead": LinearHead,
            "MaskedCrossEntropyLoss": MaskedCrossEntropyLoss,
            "MaskedTransformerEncoder": MaskedTransformerEncoder,
        },
    }
    for mod_name, attrs in targets.items():
        try:
            mod = importlib.import_module(mod_name)
        except Exception:      # the reference (or one of its dependencies) is not importable here
            continue
        for name, obj in attrs.items():
            setattr(mod, name, obj)
            patched.append(f"{mod_name}.{name}")
    return patched
