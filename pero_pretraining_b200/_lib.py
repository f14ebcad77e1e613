"""ctypes binding of libpero_b200.so (the C ABI declared in include/pero_b200.h).

There is no fallback: if the shared library is missing or a call fails, an exception is raised.
Nothing in this package imports ``oracle/``.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpero_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

c_i64 = ctypes.c_int64
c_sz = ctypes.c_size_t
c_vp = ctypes.c_void_p
c_int = ctypes.c_int
c_f32 = ctypes.c_float
c_f64 = ctypes.c_double

# name -> (restype, argtypes); mirrors include/pero_b200.h one to one.
SIGNATURES = {
    "pero_version": (c_int, []),
    "pero_strerror": (ctypes.c_char_p, [c_int]),
    "pero_check_device": (c_int, []),
    "pero_vq_codebook_bytes": (c_sz, [c_i64, c_i64]),
    "pero_vq_codebook_prepare": (c_int, [c_vp, c_i64, c_i64, c_vp, c_sz, c_vp]),
    "pero_vq_assign_workspace_bytes": (c_sz, [c_i64, c_i64, c_i64]),
    "pero_vq_assign": (c_int, [c_vp, c_i64, c_i64, c_int, c_i64, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp,
                               c_vp, c_sz, c_vp]),
    "pero_vq_assign_bf16": (c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp, c_vp]),
    "pero_vq_packed_init": (c_int, [c_vp, c_i64, c_vp]),
    "pero_vq_unpack": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp]),
    "pero_vq_gather_st": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_int, c_i64, c_i64, c_vp, c_vp]),
    "pero_vq_gather_st_mse_workspace_bytes": (c_sz, [c_i64, c_i64, c_int, c_i64]),
    "pero_vq_gather_st_mse": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_int, c_i64, c_i64, c_vp, c_f32, c_f32, c_vp, c_vp, c_sz,
                                      c_vp]),
    "pero_vq_ema_workspace_bytes": (c_sz, [c_i64, c_i64, c_i64]),
    "pero_vq_ema_accumulate": (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_sz, c_vp]),
    "pero_vq_ema_apply": (c_int, [c_vp, c_i64, c_i64, c_f64, c_f64, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp, c_sz,
                                  c_vp]),
    "pero_kmeans_update": (c_int, [c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "pero_vq_forward_workspace_bytes": (c_sz, [c_i64, c_i64, c_i64, c_int]),
    "pero_vq_forward": (c_int, [c_vp, c_i64, c_i64, c_int, c_i64, c_i64, c_vp, c_sz, c_vp, c_vp, c_vp, c_f64, c_f64, c_int,
                                c_vp, c_vp, c_vp, c_sz, c_vp]),
    "pero_vq_counts": (c_int, [c_vp, c_i64, c_i64, c_vp, c_vp]),
    "pero_mse_workspace_bytes": (c_sz, [c_i64]),
    "pero_mse_fwd": (c_int, [c_vp, c_vp, c_i64, c_f32, c_f32, c_vp, c_vp, c_sz, c_vp]),
    "pero_mse_bwd": (c_int, [c_vp, c_vp, c_i64, c_f32, c_vp, c_vp, c_vp, c_vp]),
    "pero_vq_st_commit_bwd": (c_int, [c_vp, c_vp, c_vp, c_i64, c_f32, c_vp, c_vp, c_vp]),
    "pero_head_bytes": (c_sz, [c_i64, c_i64]),
    "pero_head_prepare": (c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_sz, c_vp]),
    "pero_masked_ce_workspace_bytes": (c_sz, [c_i64, c_i64, c_i64, c_i64]),
    "pero_masked_ce_gather": (c_int, [c_vp, c_int, c_i64, c_i64, c_vp, c_i64, c_i64, c_vp, c_sz, c_vp]),
    "pero_masked_ce_fwd": (c_int, [c_vp, c_int, c_i64, c_i64, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp,
                                   c_sz, c_vp]),
    "pero_masked_ce_loss": (c_int, [c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "pero_masked_ce_bwd": (c_int, [c_vp, c_int, c_i64, c_i64, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_vp, c_f32,
                                   c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "pero_masked_ce_eval": (c_int, [c_vp, c_int, c_i64, c_i64, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_int, c_vp, c_vp, c_vp,
                                    c_vp, c_vp, c_sz, c_vp]),
    "pero_masked_ce_bwd_range": (c_int, [c_vp, c_int, c_i64, c_i64, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_vp, c_f32,
                                         c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "pero_ce_logits_fwd": (c_int, [c_vp, c_int, c_i64, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "pero_ce_logits_bwd": (c_int, [c_vp, c_int, c_i64, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_f32, c_int, c_vp, c_vp]),
    "pero_mask_compact_workspace_bytes": (c_sz, [c_i64]),
    "pero_mask_compact": (c_int, [c_vp, c_int, c_int, c_vp, c_i64, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "pero_mask_pixels": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "pero_head_argmax_prepare": (c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_sz, c_vp]),
    "pero_peer_allreduce_sum_f32": (c_int, [c_vp, c_vp, c_int, c_int, c_i64, c_i64, c_int, c_vp]),
    "pero_peer_allreduce_min_i64": (c_int, [c_vp, c_vp, c_int, c_int, c_i64, c_i64, c_int, c_vp]),
    "pero_peer_allreduce_emulate": (c_int, [c_vp, c_int, c_int, c_i64, c_i64, c_int, c_vp]),
    "pero_proj_workspace_bytes": (c_sz, [c_i64, c_i64, c_i64]),
    "pero_proj_forward": (c_int, [c_vp, c_i64, c_i64, c_int, c_i64, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "pero_gather_rows_cf": (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "pero_gemm_tn_bf16": (c_int, [c_vp, c_i64, c_vp, c_i64, c_i64, c_int, c_int, c_vp, c_vp]),
}
# present only in a dev build of the library (make DEV=1): bound when the symbol exists
DEV_SIGNATURES = {
    "pero_debug_set_timeline": (c_int, [c_vp, c_int]),
}


class PeroError(RuntimeError):
    pass


_lib = None


def build(verbose=False, dev=False):
    """Compile libpero_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).
    dev=True: the measurement build (make DEV=1: PERO_* environment knobs and timeline hooks compiled in)."""
    res = subprocess.run(["make", "-C", CSRC_DIR, "-j", "8"] + (["DEV=1"] if dev else []), capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise PeroError("building libpero_b200.so failed (see output above)")
    return LIB_PATH


def lib():
    """The loaded shared library with argtypes/restypes set.  Raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PeroError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C pero_pretraining_b200/csrc`). There is no CPU or PyTorch fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        for name, (res, args) in DEV_SIGNATURES.items():
            fn = getattr(handle, name, None)
            if fn is not None:
                fn.restype = res
                fn.argtypes = args
        _lib = handle
    return _lib


def check(code, what=""):
    if code != 0:
        msg = lib().pero_strerror(int(code)).decode()
        raise PeroError(f"{what or 'pero call'} failed with code {code}: {msg}")
