"""ctypes binding of libhwer_b200.so (the C ABI declared in include/hwer_b200.h).

There is no fallback of any kind: if the library is missing and cannot be built, or a call fails, this raises.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint32, c_void_p

from . import build as _build

HWER_OK = 0
HWER_E_INVALID = -1
HWER_E_CUDA = -2
HWER_E_ARCH = -3
HWER_E_K_TOO_LARGE = -4
HWER_E_OVERFLOW = -5
HWER_E_NOMEM = -6
HWER_E_PEER = -7
IPC_HANDLE_BYTES = 64
PHASE_SEARCH, PHASE_MERGE, PHASE_COLLECT, PHASE_ALL, PHASE_OWNED = 1, 2, 4, 7, 8
MODE_EXACT = 0
MODE_BF16 = 1
SCORE_PAIR, SCORE_DIST, SCORE_GIVEN, SCORE_EUCLID = 0, 1, 2, 3

# name -> (restype, argtypes); kept in one table so tests can check every symbol of the header is exported
SIGNATURES = {
    "hwer_last_error": (c_char_p, []),
    "hwer_version": (c_int, []),
    "hwer_shadow_width": (c_int32, [c_int32]),
    "hwer_blend_normalize": (c_int, [c_void_p, c_void_p, c_float, c_void_p, c_int64, c_int32, c_void_p, c_void_p,
                                     c_int32, c_void_p]),
    "hwer_make_shadow": (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_int32, c_void_p]),
    "hwer_norm_stats": (c_int, [c_void_p, c_int64, c_int32, c_float, c_void_p, c_void_p]),
    "hwer_index_create": (c_int, [POINTER(c_void_p), c_void_p, c_void_p, c_int64, c_int32, c_int32, c_float, c_int32]),
    "hwer_index_destroy": (c_int, [c_void_p]),
    "hwer_topk": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_uint32, c_int64, c_void_p, c_void_p,
                          c_void_p, c_void_p]),
    "hwer_topk_exhaustive": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int64, c_void_p, c_void_p, c_void_p,
                                     c_void_p]),
    "hwer_topk_finish": (c_int, [c_void_p, c_void_p, POINTER(c_uint32)]),
    "hwer_profile": (c_int, [c_void_p, c_int]),
    "hwer_profile_read": (c_int, [c_void_p, c_void_p, POINTER(c_double), POINTER(c_int64), POINTER(c_int64)]),
    "hwer_profile_stages": (c_int, [c_void_p, c_void_p, POINTER(c_double)]),
    "hwer_profile_launches": (c_int, [c_void_p, c_void_p, POINTER(c_double), c_int32, POINTER(c_int32)]),
    "hwer_debug_scores": (c_int, [c_void_p, c_void_p, c_int32, c_void_p, c_int64, c_void_p]),
    "hwer_merge_topk": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hwer_exchange_bytes": (c_int64, [c_int32, c_int32, c_int32]),
    "hwer_peer_alloc": (c_int, [c_int64, POINTER(c_void_p), c_void_p]),
    "hwer_peer_open": (c_int, [c_void_p, POINTER(c_void_p)]),
    "hwer_peer_close": (c_int, [c_void_p]),
    "hwer_peer_free": (c_int, [c_void_p]),
    "hwer_exchange_create": (c_int, [POINTER(c_void_p), c_int32, c_int32, c_int32, c_int32, POINTER(c_void_p), c_int32]),
    "hwer_exchange_configure": (c_int, [c_void_p, c_int64, c_int32]),
    "hwer_exchange_destroy": (c_int, [c_void_p]),
    "hwer_topk_sharded": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_uint32, c_int64, c_void_p,
                                  c_void_p, c_void_p, c_int32, c_void_p]),
    "hwer_exchange_error": (c_int, [c_void_p, c_void_p]),
    "hwer_pair_score": (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "hwer_compose_queries": (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_int32, c_void_p, c_void_p]),
    "hwer_average_embeddings": (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_int32, c_void_p, c_void_p]),
    "hwer_gather_rows": (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_int64, c_void_p, c_void_p]),
    "hwer_map_rows": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "hwer_rerank": (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p,
                            c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hwer_hit_rank_metrics": (c_int, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "hwer_ncf_param_count": (c_int64, [c_int32, c_int32]),
    "hwer_ncf_score": (c_int, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                               c_void_p]),
    "hwer_gcn_infer": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                               c_void_p, c_void_p]),
    "hwer_eval_metrics": (c_int, [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_int32, c_int64, c_void_p, c_void_p, c_void_p]),
    "hwer_link_metrics": (c_int, [c_void_p, c_void_p, c_int64, c_float, c_void_p, c_void_p]),
}

_lib = None


class HwerError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("hwer_b200 error %d: %s" % (code, message))
        self.code = code


def library_path():
    return os.environ.get("HWER_B200_LIB") or _build.LIB


def lib():
    """Loads (building first if the in-tree .so is missing or stale and nvcc is available) the native library."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("HWER_B200_LIB") or _build.LIB      # override: A/B runs of two builds (scripts/ab_rounds.py)
    if path == _build.LIB and _build.is_stale():
        try:
            _build.build()
        except Exception as e:
            if not os.path.exists(path):
                raise RuntimeError("libhwer_b200.so is not built and cannot be built here (%s). "
                                   "hwer_b200 has no CPU fallback." % e)
    handle = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)   # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return _lib


def check(code):
    if code != HWER_OK:
        raise HwerError(code, lib().hwer_last_error().decode())
    return code
