"""ctypes binding of libb200ivfpq.so (the C-ABI declared in include/b200_ivfpq.h).

There is no CPU path: if the shared library is missing, or no CUDA device is present when an index is
created, the call fails loudly.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200ivfpq.so")

_c_f32p = ctypes.c_void_p
_lib = None

# every symbol include/b200_ivfpq.h declares: (name, restype, argtypes)
_I, _L, _P = ctypes.c_int, ctypes.c_int64, ctypes.c_void_p
SYMBOLS = [
    ("b200_ivfpq_last_error", ctypes.c_char_p, []),
    ("b200_ivfpq_version", ctypes.c_char_p, []),
    ("b200_ivfpq_launch_count", _L, []),
    ("b200_ivfpq_create", _I, [_I, _L, _I, _I, ctypes.POINTER(_P)]),
    ("b200_ivfpq_destroy", _I, [_P]),
    ("b200_ivfpq_set_codebooks", _I, [_P, _P, _P]),
    ("b200_ivfpq_set_lists", _I, [_P, _P, _P, _P, _L]),
    ("b200_ivfpq_coarse", _I, [_P, _L, _P, _I, _P, _P, _P]),
    ("b200_ivfpq_coarse_scores", _I, [_P, _L, _P, _P, _P]),
    ("b200_ivfpq_coarse_fallbacks", _I, [_P, ctypes.POINTER(_L)]),
    ("b200_ivfpq_search", _I, [_P, _L, _P, _I, _I, _P, _P, _P]),
    ("b200_ivfpq_search_preassigned", _I, [_P, _L, _P, _I, _I, _P, _P, _P, _P]),
    ("b200_ivfpq_prepare_queries", _I, [_P, _L, _P, _P]),
    ("b200_ivfpq_search_preassigned_begin", _I, [_P, _L, _P, _I, _I, _P, _L, _L, _P, _P]),
    ("b200_ivfpq_search_preassigned_finish", _I, [_P, _P, _P, _P]),
    ("b200_ivfpq_search_host", _I, [_P, _L, _P, _I, _I, _P, _P]),
    ("b200_ivfpq_assign_encode", _I, [_P, _L, _P, _P, _P, _P]),
    ("b200_ivfpq_segment_sums", _I, [_L, _I, _P, _P, _P, _P, _P]),
    ("b200_ivfpq_merge_shards", _I, [_I, _L, _I, _P, _P, _P, _P, _P]),
    ("b200_ivfpq_merge_shards_peer", _I, [_I, _L, _I, _P, _L, _L, _P, _P, _P]),
    ("b200_ivfpq_set_stage_timing", _I, [_P, _I]),
    ("b200_ivfpq_get_stage_ms", _I, [_P, _P]),
    ("b200_ivfpq_get_last_scan_stats", _I, [_P, _P, _P]),
    ("b200_ivfpq_get_filter_stats", _I, [_P, _P, _I]),
    ("b200_ivfpq_get_filter_ms", _I, [_P, _P]),
]


def load():
    """Load the CUDA library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the B200 CUDA extension is not built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or csrc/build.sh). "
            "b200ivfpq has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int) -> None:
    """Faiss raises RuntimeError from FAISS_THROW; mirror that from the C return code."""
    if rc != 0:
        msg = load().b200_ivfpq_last_error()
        raise RuntimeError(f"b200ivfpq error {rc}: {msg.decode() if msg else '?'}")


def launch_count() -> int:
    return int(load().b200_ivfpq_launch_count())
