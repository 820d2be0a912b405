"""ctypes binding of librir.so — the C ABI declared in include/rir.h.

There is deliberately no fallback: if the library is missing or the device is not a B200 the calls raise.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint8, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "librir.so")

# constants mirrored from include/rir.h
RIR_OK = 0
RIR_E_ARG, RIR_E_ARCH, RIR_E_CUDA, RIR_E_WORKSPACE = -1, -2, -3, -4
RIR_F32, RIR_BF16, RIR_FP8E4M3 = 0, 1, 2
RIR_POOL_GEM, RIR_POOL_MAX, RIR_POOL_AVG = 0, 1, 2
RIR_PATH_AUTO, RIR_PATH_STREAM, RIR_PATH_MMA, RIR_PATH_EXACT = 0, 1, 2, 3
RIR_WS_CLEAN = 0x100
RIR_EXCHANGE_ASYNC = 0x200
RIR_MAP_OK, RIR_MAP_EMPTY_OK, RIR_MAP_NO_POS_RETRIEVED = 0, 1, 2

PATHS = {"auto": RIR_PATH_AUTO, "stream": RIR_PATH_STREAM, "mma": RIR_PATH_MMA, "exact": RIR_PATH_EXACT}

# every symbol include/rir.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "rir_version": (c_int, []),
    "rir_last_error": (c_char_p, []),
    "rir_device_check": (c_int, []),
    "rir_pool": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_float, c_void_p,
                         c_void_p]),
    "rir_l2_normalize": (c_int, [c_void_p, c_int64, c_int, c_float, c_void_p, c_void_p]),
    "rir_whiten": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "rir_whiten_prepared_bytes": (c_size_t, [c_int, c_int]),
    "rir_whiten_prepare": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "rir_gem_l2_whiten_workspace": (c_size_t, [c_int, c_int, c_int]),
    "rir_gem_l2_whiten": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_float, c_void_p,
                                  c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "rir_whiten_prepared": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                    c_size_t, c_void_p]),
    "rir_pca_covariance_workspace": (c_size_t, [c_int64, c_int]),
    "rir_pca_covariance": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "rir_scale_mean_l2": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "rir_pack_descriptors": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "rir_sim_topk_workspace": (c_size_t, [c_int, c_int64, c_int, c_int, c_int]),
    "rir_sim_topk": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_int64,
                             c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "rir_sim_topk_workspace_init": (c_int, [c_void_p, c_size_t, c_void_p]),
    "rir_profile_scan_begin": (c_int, []),
    "rir_profile_scan_end": (c_int, [POINTER(c_float), c_int, POINTER(c_int)]),
    "rir_profile_scan_pause": (c_int, [c_int]),
    "rir_profile_timeline": (c_int, [c_void_p, c_int]),
    "rir_rescore_topk": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_int64, c_int, c_void_p,
                                 c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "rir_merge_topk_workspace": (c_size_t, [c_int, c_int, c_int]),
    "rir_merge_topk": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                               c_void_p]),
    "rir_exchange_bytes": (c_size_t, [c_int, c_int, c_int]),
    "rir_peer_alloc": (c_int, [c_size_t, POINTER(c_void_p)]),
    "rir_peer_free": (c_int, [c_void_p]),
    "rir_peer_export": (c_int, [c_void_p, c_void_p]),
    "rir_peer_open": (c_int, [c_void_p, POINTER(c_void_p)]),
    "rir_peer_close": (c_int, [c_void_p]),
    "rir_sim_topk_sharded": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_int, c_int,
                                     c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p, c_int, c_int,
                                     ctypes.c_uint32, c_int, c_int, POINTER(c_void_p)]),
    "rir_exchange_join": (c_int, [c_void_p, ctypes.c_uint32, c_void_p]),
    "rir_exchange_sync": (c_int, [c_void_p, ctypes.c_uint32]),
    "rir_search_host_workspace": (c_size_t, [c_int, c_int64, c_int, c_int, c_int]),
    "rir_search_host": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int64, c_int, c_int, c_int64, c_void_p,
                                c_void_p, c_void_p, c_size_t, c_int, c_void_p, c_int, c_int, ctypes.c_uint32, c_int, c_int,
                                POINTER(c_void_p)]),
    "rir_aqe_accumulate": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_int,
                                   c_int, c_int, c_float, c_void_p, c_void_p]),
    "rir_aqe_finalize": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                 c_void_p]),
    "rir_compute_map": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, POINTER(c_int32), c_int, POINTER(c_int32), c_int, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p]),
    "rir_compute_map_at": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, POINTER(c_int32), c_int, POINTER(c_int32), c_int, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    "rir_gnd_scores": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_int, c_int64, c_void_p,
                               c_void_p, c_int64, c_void_p, c_void_p]),
    "rir_rank_thresholds": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "rir_rank_count": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_int, c_int64, c_void_p,
                               c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
}

_lib = None


class RirError(RuntimeError):
    """A librir.so entry point returned a negative RIR_E_* code."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"librir error {code}: {msg}")
        self.code = code


def load() -> ctypes.CDLL:
    """dlopen librir.so (built by research_image_retrieval_b200.build) and type every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: run `python -m research_image_retrieval_b200.build` (needs nvcc). "
            "This package has no CPU / PyTorch fallback for the retrieval hot path."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError == ABI mismatch with include/rir.h
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != RIR_OK:
        msg = load().rir_last_error()
        raise RirError(rc, msg.decode("utf-8", "replace") if msg else "")


def stream_ptr() -> int:
    """cudaStream_t of torch's current stream, so launches are ordered with the caller's torch work."""
    import torch

    return torch.cuda.current_stream().cuda_stream


def require_cuda_tensor(t, name: str, dtype=None):
    import torch

    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA tensor (this path has no CPU implementation)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must have dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t
