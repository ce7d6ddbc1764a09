"""ctypes binding of libalscore.so (include/alscore.h).  No fallback: if the CUDA library is
missing or there is no B200, every entry point raises."""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
# ALS_LIB_TAG=<tag> loads libalscore_<tag>.so: a bring-up build (trace instrumentation, single class count) made with
# ALS_BUILD_TAG=<tag> NVCC_EXTRA=... python -m semanticsegmentationactivelearning_b200.build -- never the shipped library
_TAG = os.environ.get("ALS_LIB_TAG", "")
LIB_PATH = os.path.join(PKG, "libalscore%s.so" % ("_" + _TAG if _TAG else ""))

ALS_OK = 0
ALS_ERR_INVALID = -1
ALS_ERR_UNSUPPORTED = -2
ALS_ERR_CUDA = -3
ALS_ERR_NOMEM = -4
ALS_ERR_STATE = -5

ALS_F32, ALS_BF16 = 0, 1
ALS_ENTROPY, ALS_MARGIN, ALS_CONFIDENCE, ALS_VARIANCE = 0, 1, 2, 3

_i64 = C.c_int64
_p = C.c_void_p

# name -> (restype, argtypes); must list every symbol include/alscore.h declares
SIGNATURES = {
    "als_version": (C.c_int, []),
    "als_device_count": (C.c_int, []),
    "als_ctx_create": (C.c_int, [C.c_int, C.POINTER(_p)]),
    "als_ctx_destroy": (C.c_int, [_p]),
    "als_ctx_set_stream": (C.c_int, [_p, _p]),
    "als_last_error": (C.c_char_p, [_p]),
    "als_measure_from_name": (C.c_int, [C.c_char_p, C.POINTER(C.c_int)]),
    "als_launch_count": (_i64, [_p]),
    "als_ctx_enable_timing": (C.c_int, [_p, C.c_int]),
    "als_last_scoring_ms": (C.c_int, [_p, C.POINTER(C.c_float)]),
    "als_score": (C.c_int, [_p, _p, C.c_int, _i64, _i64, _i64, _i64, _i64, C.c_int, _p, _p, _p, _p, C.c_float, _p]),
    "als_score_host": (C.c_int, [_p, _p, C.c_int, _i64, _i64, _i64, _i64, _i64, C.c_int, _p, _p, _p, _p, C.c_float]),
    "als_score_dlpack": (C.c_int, [_p, _p, C.c_int, _p, _p]),
    "als_head_prepare": (C.c_int, [_p, _p, _i64]),
    "als_head_geometry": (C.c_int, [_i64, _p, _p]),
    "als_head_pack_weights": (C.c_int, [_p, _i64, _p, _i64]),
    "als_head_supported": (C.c_int, [_i64, C.c_int, _i64]),
    "als_score_features": (C.c_int, [_p, _p, _i64, _i64, _i64, _i64, C.c_int, _p, _p, _p, _p, C.c_float, _p]),
    "als_pool_score_features_batch": (C.c_int, [_p, _p, C.c_int, _i64, _i64, _i64, _i64, C.c_int, _p]),
    "als_pool_begin": (C.c_int, [_p, _i64]),
    "als_pool_score_batch": (C.c_int, [_p, _p, C.c_int, C.c_int, _i64, _i64, _i64, _i64, _i64, C.c_int, _p]),
    "als_pool_scores": (C.c_int, [_p, _p, _i64]),
    "als_pool_select": (C.c_int, [_p, _p, _i64, _i64, _p, _p, C.POINTER(_i64)]),
    "als_rank_pool": (C.c_int, [_p, _p, C.c_int, C.c_int, _i64, _i64, _i64, _i64, _i64, C.c_int, _p, _i64, _p, _i64, _i64,
                                _p, _p, C.POINTER(_i64)]),
    "als_select_smallest": (C.c_int, [_p, _p, _p, _i64, _i64, _p, _p, _p]),
    "als_comm_unique_id": (C.c_int, [_p]),
    "als_comm_init_rank": (C.c_int, [_p, C.c_int, C.c_int, _p]),
    "als_comm_init_all": (C.c_int, [C.POINTER(_p), C.c_int]),
    "als_comm_destroy": (C.c_int, [_p]),
    "als_pool_select_global": (C.c_int, [_p, _p, _i64, _i64, _i64, _i64, _i64, _p, _p, C.POINTER(_i64)]),
    "als_pool_select_global_all": (C.c_int, [C.POINTER(_p), C.c_int, _p, _i64, _i64, _p, _p, _i64, _p, _p,
                                             C.POINTER(_i64)]),
    "als_mc_begin": (C.c_int, [_p, C.c_int, _i64, _i64, _i64, _i64, _p]),
    "als_mc_add_sample": (C.c_int, [_p, _p, C.c_int]),
    "als_mc_finish": (C.c_int, [_p, C.c_int, _p, _p, _p, _p, C.c_float]),
    "als_synth_logits": (C.c_int, [_p, _p, C.c_int, _i64, _i64, _i64, _i64, _i64, _i64, C.c_uint64, C.c_int, _p]),
    "als_flush_l2": (C.c_int, [_p, _p]),
    "als_describe_launch": (C.c_int, [_p, C.c_int, _i64, _i64, _i64, _i64, _i64, C.c_int, C.c_char_p,
                                      C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                      C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "als_describe_head_launch": (C.c_int, [_p, _i64, C.c_int, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                           C.POINTER(C.c_int)]),
}

ALS_VERSION = 110
# `stream` argument meaning "the context's stream" (include/alscore.h: ALS_STREAM_CTX); 0 / None is the legacy default stream
ALS_STREAM_CTX = C.c_void_p(-1)

_lib = None


class AlscoreUnavailable(RuntimeError):
    """The CUDA library is not built / not loadable.  There is deliberately no CPU path."""


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AlscoreUnavailable(
            "%s not found: build it with `python -m semanticsegmentationactivelearning_b200.build` "
            "(needs nvcc; the pool-scoring path has no CPU fallback)" % LIB_PATH)
    _point_at_nccl()
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise AlscoreUnavailable("cannot load %s: %s" % (LIB_PATH, e)) from e
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _point_at_nccl() -> None:
    """The multi-GPU exchange binds NCCL at run time (csrc/comm.cu): a copy the process already holds wins, else
    ALS_NCCL_LIB, else the loader path.  Default ALS_NCCL_LIB to the wheel torch ships (nvidia-nccl-cu12)."""
    if os.environ.get("ALS_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for root in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(root, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["ALS_NCCL_LIB"] = cand
                return
    except Exception:  # pragma: no cover
        pass


def last_error(ctx=None) -> str:
    msg = load().als_last_error(ctx)
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, ctx=None) -> None:
    """Map als_status to the exception types the reference raises at the same places."""
    if rc == ALS_OK:
        return
    msg = last_error(ctx) or last_error(None) or ("alscore error %d" % rc)
    if rc == ALS_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)      # active_learning.py:259-260
    if rc == ALS_ERR_INVALID:
        raise ValueError(msg)
    if rc == ALS_ERR_NOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)
