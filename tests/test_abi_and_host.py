"""CPU-only checks: the C-ABI library loads and exports every symbol include/pero_b200.h declares, host-side
size queries / error paths behave (no compute call needs a GPU), and the host logic of the Python mirror
(mask compaction on the host, shard bounds, the packed (distance, index) order) is right."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(dev_only=False):
    """Functions declared in the header; the `#ifdef PERO_DEV_BUILD` block (measurement hooks that the production
    library does not export) is returned separately."""
    text = open(os.path.join(ROOT, "include", "pero_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    dev_blocks = re.findall(r"#ifdef PERO_DEV_BUILD(.*?)#endif", text, flags=re.S)
    if dev_only:
        text = "\n".join(dev_blocks)
    else:
        text = re.sub(r"#ifdef PERO_DEV_BUILD.*?#endif", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pero_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from pero_pretraining_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.lib()


def test_library_exports_every_declared_symbol(lib):
    from pero_pretraining_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 25
    raw = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared if not hasattr(raw, s)]
    assert not missing, f"declared in include/pero_b200.h but not exported: {missing}"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes signatures must mirror the header one to one"
    dev = _declared_symbols(dev_only=True)
    assert sorted(_lib.DEV_SIGNATURES) == dev and dev
    exported_dev = [s for s in dev if hasattr(raw, s)]
    # production build (what __graft_entry__.build() makes): no measurement hooks, no environment knobs
    assert exported_dev in ([], dev), "either a production build (no dev symbols) or a full dev build"
    if not exported_dev:
        blob = open(_lib.LIB_PATH, "rb").read()
        assert b"PERO_GEMM_MAX_CTAS" not in blob and b"PERO_PDL" not in blob, "production build must not read tuning knobs"


def test_version_strerror_and_size_queries(lib):
    assert lib.pero_version() >= 100
    assert lib.pero_strerror(0) == b"success"
    for code in (-1, -2, -3, -4, -5, -6, -7):
        assert len(lib.pero_strerror(code)) > 8
    # prepared codebook: bf16 [K, Dp] + |c|^2 [Kp]; Dp, Kp padded to 64 / 256
    assert lib.pero_vq_codebook_bytes(8192, 256) == 8192 * 256 * 2 + 8192 * 4
    assert lib.pero_vq_codebook_bytes(1000, 200) >= 1000 * 256 * 2 + 1024 * 4
    assert lib.pero_vq_codebook_bytes(0, 5) == 0
    assert lib.pero_vq_assign_workspace_bytes(8192, 8192, 256) >= 8192 * 256 * 2 + 8192 * 8
    assert lib.pero_vq_ema_workspace_bytes(8192, 8192, 256) > 4 * 8192 * 4
    assert lib.pero_head_bytes(4096, 512) == 4096 * 512 * 2 + 4096 * 4      # ONE bf16 copy of W + the bias
    assert lib.pero_masked_ce_workspace_bytes(1024, 154, 4096, 512) > 0
    assert lib.pero_mse_workspace_bytes(10) >= 256 and lib.pero_mask_compact_workspace_bytes(1000) >= 256


def test_argument_errors_do_not_need_a_gpu(lib):
    assert lib.pero_vq_codebook_prepare(None, 8, 8, None, 0, None) == -5
    assert lib.pero_vq_assign(None, 1, 1, 0, 8, 8, None, 0, None, None, None, None, None, 0, None) == -5
    assert lib.pero_vq_assign(None, -1, 1, 0, 8, 8, None, 0, None, None, None, None, None, 0, None) == -1
    assert lib.pero_vq_assign(None, 0, 128, 1, 8, 8, None, 0, None, None, None, None, None, 0, None) == 0    # zero frames: no-op
    assert lib.pero_mse_fwd(None, None, 4, 1.0, 0.0, None, None, 0, None) == -5
    assert lib.pero_masked_ce_fwd(None, 0, 8, 8, None, 0, None, None, 8, None, None, None, 0, None) == -1   # empty mask
    assert lib.pero_vq_ema_apply(None, 8, 8, 0.99, 1e-5, None, None, None, None, 0, None, 0, None) == -5


def test_python_mirror_fails_loudly_without_gpu():
    from pero_pretraining_b200 import KMeansLabeller, LinearHead, MaskedCrossEntropyLoss, PeroError, VectorQuantizer
    with pytest.raises(PeroError):
        VectorQuantizer(8, 4, 0.25, 0.99)(torch.randn(1, 4, 1, 3))
    with pytest.raises(PeroError):
        KMeansLabeller(torch.randn(4, 4))
    with pytest.raises(PeroError):
        LinearHead(4, 8).masked_loss(torch.randn(1, 3, 4), torch.zeros(1, 3, dtype=torch.long), np.ones((1, 3), dtype=int))
    with pytest.raises(PeroError):
        MaskedCrossEntropyLoss()(torch.randn(1, 3, 8), torch.zeros(1, 3, dtype=torch.long), torch.ones(1, 3, dtype=torch.long))
    # VQVAE.quantize / labels: whichever path 'auto' takes, there is no CPU quantizer behind it
    from pero_pretraining_b200 import VQVAE, ops

    class _Pass(torch.nn.Module):
        out_channels = base_channels = 6

        def forward(self, x):
            return x

    m = VQVAE(_Pass(), _Pass(), 8, 4)
    assert m.fuse_projections == 'auto'
    for fuse in ('auto', True, False):
        m.fuse_projections = fuse
        with pytest.raises(PeroError):
            m.quantize(torch.randn(1, 6, 1, 3))
        with pytest.raises(PeroError), torch.no_grad():
            m.labels(torch.randn(1, 6, 1, 3))
    with pytest.raises((PeroError, TypeError)):
        ops.proj_forward(torch.randn(3, 6), torch.randn(4, 6), None, 3, 1, False)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pero_pretraining_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                hits = [ln for ln in src.splitlines() if re.search(r"^\s*(from|import)\s+\S*oracle|oracle[./]pero_oracle|import_module\(.*oracle", ln)]
                assert not hits, f"{f} imports the oracle: {hits}"


def test_module_surface_matches_reference_signatures():
    """Names, constructor arguments and state_dict keys of SURVEY §8b."""
    import inspect
    from pero_pretraining_b200 import LinearHead, MaskedCrossEntropyLoss, MaskedTransformerEncoder, VQVAE, VectorQuantizer
    assert list(inspect.signature(VectorQuantizer.__init__).parameters)[1:] == \
        ["num_embeddings", "embeddings_dim", "commitment_cost", "decay", "epsilon"]
    assert inspect.signature(VectorQuantizer.__init__).parameters["epsilon"].default == 1e-5
    p = inspect.signature(VQVAE.__init__).parameters
    assert list(p)[1:] == ["encoder", "decoder", "num_embeddings", "embeddings_dim", "commitment_cost", "decay", "reconstruction_loss"]
    assert (p["commitment_cost"].default, p["decay"].default, p["reconstruction_loss"].default) == (0.25, 0.99, "mse")
    vq = VectorQuantizer(16, 8, 0.25, 0.99)
    assert set(vq.state_dict()) == {"embedding.weight", "ema_w", "ema_cluster_size"}
    assert isinstance(vq.ema_w, torch.nn.Parameter) and float(vq.ema_cluster_size.sum()) == 0.0
    for attr in ("num_embeddings", "embeddings_dim", "commitment_cost", "decay", "epsilon"):
        assert hasattr(vq, attr)
    w0 = VectorQuantizer(16, 8, 0.25, 0.0).embedding.weight
    assert float(w0.abs().max()) <= 1 / 16                      # U(-1/K, 1/K) when decay == 0 (:190)
    head = LinearHead()
    assert (head.linear.in_features, head.linear.out_features) == (512, 4096)
    assert set(head.state_dict()) == {"linear.weight", "linear.bias"}
    assert MaskedCrossEntropyLoss().unmasked_weight is None
    assert list(inspect.signature(MaskedTransformerEncoder.forward).parameters)[1:] == ["x", "labels", "mask"]


def test_host_mask_compaction_matches_boolean_indexing():
    from pero_pretraining_b200.masked_pretraining import _rows_from_mask
    rng = np.random.default_rng(3)
    labels = torch.from_numpy(rng.integers(-1, 6, size=(4, 50)))
    mask = (rng.random((4, 50)) < 0.3).astype(int) * (labels.numpy() >= 0)
    rows, m = _rows_from_mask(mask, labels, 1, False, torch.device("cpu"))
    assert m == mask.sum() and np.array_equal(rows.numpy(), np.flatnonzero(mask.reshape(-1) == 1))
    rows0, m0 = _rows_from_mask(mask, labels, 0, True, torch.device("cpu"))
    ref0 = np.flatnonzero((mask.reshape(-1) == 0) & (labels.numpy().reshape(-1) >= 0))
    assert m0 == ref0.size and np.array_equal(rows0.numpy(), ref0) and rows0.dtype == torch.int32


def test_shard_bounds_cover_everything_once():
    from pero_pretraining_b200.sharding import shard_bounds
    for total, world in ((65536, 8), (1000, 3), (5, 8), (0, 2)):
        spans = [shard_bounds(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_packed_distance_index_order():
    """Signed int64 order of (order_key(d) << 32 | idx) == lexicographic (distance, index) order, for
    negative, zero and positive distances (|c|^2 - 2<x,c> can be negative)."""
    from pero_pretraining_b200.sharding import pack_dist_index_reference, unpack_index_reference
    d = torch.tensor([-3.5, -0.0, 0.0, 1e-30, 2.0, 2.0, float("inf"), -1e20, 7.25])
    idx = torch.tensor([5, 1, 0, 9, 4, 3, 0, 2, 4294967295 // 2])
    packed = pack_dist_index_reference(d, idx)
    order = torch.argsort(packed)
    pairs = [(float(d[i]), int(idx[i])) for i in order]
    ref = sorted(zip(d.tolist(), idx.tolist()), key=lambda t: (t[0], t[1]))
    # -0.0 and 0.0 are distinct keys (-0.0 first); Python's sort treats them as equal, so compare canonically
    assert [(p[0] + 0.0, p[1]) for p in pairs if p[0] != 0.0] == [(p[0] + 0.0, p[1]) for p in ref if p[0] != 0.0]
    assert torch.equal(unpack_index_reference(packed), idx)
    assert int(packed.max()) < 0x7FFFFFFFFFFFFFFF           # INT64_MAX is reserved for "empty"


def test_label_wire_format_roundtrip(tmp_path):
    """"<line_id> l0 l1 ...\\n" with frames filtered by image_mask == 1 (produce_kmeans_labels.py:83-85,
    scripts/common.py:51-54) — host-side code, no device involved."""
    import os
    import numpy as np
    from pero_pretraining_b200.labels_io import LabelWriter, format_label_line, load_labels, save_labels
    assert format_label_line("a.jpg", [3, 0, 12]) == "a.jpg 3 0 12\n"
    assert format_label_line("empty.jpg", []) == "empty.jpg \n"           # what the reference's f-string writes
    p = os.path.join(tmp_path, "l.txt")
    with LabelWriter(p) as w:
        w.write_batch(["x", "y"], np.array([[1, 1, 0, 0], [1, 0, 1, 1]]), np.array([[5, 6, 7, 8], [9, 10, 11, 12]]))
    assert open(p).read() == "x 5 6\ny 9 11 12\n"
    assert load_labels(p) == {"x": [5, 6], "y": [9, 11, 12]}
    q = os.path.join(tmp_path, "m.txt")
    save_labels({"x": [5, 6], "y": [9, 11, 12]}, q)
    assert open(q).read() == open(p).read()


def test_create_mask_matches_reference_batch_operator():
    """masked_pretraining/batch_operator.py:27-32 restated: same global-numpy seed -> same mask; padding (-1) never masked."""
    import numpy as np
    from pero_pretraining_b200 import create_mask
    labels = np.random.default_rng(0).integers(-1, 50, size=(6, 40))
    np.random.seed(123)
    want = (np.random.rand(*labels.shape) < 0.15).astype(int) * (labels >= 0).astype(int)      # the reference's two lines
    np.random.seed(123)
    got = create_mask(labels, 0.15)
    assert got.dtype == want.dtype and np.array_equal(got, want)
    assert not got[labels < 0].any()
    g = create_mask(labels, 0.5, np.random.default_rng(1))
    assert g.shape == labels.shape and set(np.unique(g)) <= {0, 1}


def test_header_is_plain_c(tmp_path):
    """include/pero_b200.h is the drop-in boundary: it must compile as C99 (no C++-isms, no CUDA or torch types), so
    that cgo / JNI / ctypes-style bindings of any host language can consume it."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "use_header.c"
    src.write_text('#include "pero_b200.h"\n'
                   'int probe(void) { return pero_version() + (int)pero_vq_codebook_bytes(8, 8) + PERO_PEER_HEADER_BYTES; }\n')
    res = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(root, "include"),
                          str(src)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
