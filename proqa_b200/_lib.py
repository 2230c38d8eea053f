"""ctypes binding of libproqa_b200.so (the C ABI declared in include/proqa_b200.h).

The library is loaded lazily on first use and NEVER at import of ``faiss``: the reference forks its worker
pool after ``import faiss`` (retrieval/eval_retrieval.py:4,92-96).  Loading the .so does not touch CUDA either
(the C side initialises the device on the first add/search).  If the library is missing the call fails
loudly — there is no Python/CPU fallback.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

PQ_OK = 0
PQ_MAX_K = 15360
TIER_AUTO, TIER_FP32, TIER_BF16 = 0, 1, 2

class KMeansParams(ctypes.Structure):
    """struct pq_kmeans_params (include/proqa_b200.h)."""
    _fields_ = [("niter", ctypes.c_int), ("nredo", ctypes.c_int), ("verbose", ctypes.c_int), ("spherical", ctypes.c_int),
                ("min_points_per_centroid", ctypes.c_int), ("max_points_per_centroid", ctypes.c_int), ("seed", ctypes.c_int64)]


# name -> (restype, argtypes); the single source the symbol test checks against include/proqa_b200.h
_f32p = ctypes.POINTER(ctypes.c_float)
_i64p = ctypes.POINTER(ctypes.c_int64)
_vp = ctypes.c_void_p
SIGNATURES = {
    "pq_index_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_vp)]),
    "pq_index_free": (None, [_vp]),
    "pq_index_add": (ctypes.c_int, [_vp, ctypes.c_int64, _vp]),
    "pq_index_add_device": (ctypes.c_int, [_vp, ctypes.c_int64, _vp]),
    "pq_index_add_f16": (ctypes.c_int, [_vp, ctypes.c_int64, _vp]),
    "pq_index_search": (ctypes.c_int, [_vp, ctypes.c_int64, _vp, ctypes.c_int64, _vp, _vp]),
    "pq_index_search_device": (ctypes.c_int, [_vp, ctypes.c_int64, _vp, ctypes.c_int64, _vp, _vp]),
    "pq_index_reset": (ctypes.c_int, [_vp]),
    "pq_index_ntotal": (ctypes.c_int64, [_vp]),
    "pq_index_d": (ctypes.c_int, [_vp]),
    "pq_index_metric": (ctypes.c_int, [_vp]),
    "pq_index_set_id_base": (ctypes.c_int, [_vp, ctypes.c_int64]),
    "pq_index_set_tier": (ctypes.c_int, [_vp, ctypes.c_int]),
    "pq_index_set_stream": (ctypes.c_int, [_vp, _vp, ctypes.c_int]),
    "pq_index_set_profile": (ctypes.c_int, [_vp, ctypes.c_int]),
    "pq_index_last_stats": (ctypes.c_int, [_vp, _i64p, ctypes.c_int]),
    "pq_merge_shard_results": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, _vp, _vp,
                                              _vp, _vp]),
    "pq_merge_shard_results_async": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, _vp, _vp,
                                                    _vp, _vp, _vp]),
    "pq_index_share_alloc": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.POINTER(_vp), _i64p]),
    "pq_index_share_connect": (ctypes.c_int, [_vp, ctypes.POINTER(_vp)]),
    "pq_index_share_begin": (ctypes.c_int, [_vp, ctypes.c_uint32]),
    "pq_index_share_close": (ctypes.c_int, [_vp]),
    "pq_index_get_bound_scalars": (ctypes.c_int, [_vp, _f32p]),
    "pq_index_set_bound_scalars": (ctypes.c_int, [_vp, ctypes.c_float, ctypes.c_float]),
    "pq_ipc_export": (ctypes.c_int, [_vp, _vp]),
    "pq_ipc_open": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.POINTER(_vp)]),
    "pq_ipc_close": (ctypes.c_int, [ctypes.c_int, _vp]),
    "pq_enable_peer_access": (ctypes.c_int, [ctypes.c_int, ctypes.c_int]),
    "pq_multi_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(_vp)]),
    "pq_multi_free": (None, [_vp]),
    "pq_multi_add": (ctypes.c_int, [_vp, ctypes.c_int64, _vp]),
    "pq_multi_search": (ctypes.c_int, [_vp, ctypes.c_int64, _vp, ctypes.c_int64, _vp, _vp]),
    "pq_multi_reset": (ctypes.c_int, [_vp]),
    "pq_multi_ntotal": (ctypes.c_int64, [_vp]),
    "pq_multi_n_devices": (ctypes.c_int, [_vp]),
    "pq_multi_mode": (ctypes.c_int, [_vp]),
    "pq_multi_last_stats": (ctypes.c_int, [_vp, _i64p, ctypes.c_int]),
    "pq_multi_first_shard": (_vp, [_vp]),
    "pq_xchg_bytes_needed": (ctypes.c_int64, [ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int]),
    "pq_xchg_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.POINTER(_vp), ctypes.POINTER(_vp), _i64p]),
    "pq_xchg_connect": (ctypes.c_int, [_vp, ctypes.POINTER(_vp)]),
    "pq_xchg_run": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, _vp, _vp, _vp, _vp, ctypes.c_uint64, _vp]),
    "pq_xchg_check": (ctypes.c_int, [_vp]),
    "pq_xchg_free": (None, [_vp]),
    "pq_kmeans_default_params": (None, [ctypes.POINTER(KMeansParams)]),
    "pq_kmeans_train": (ctypes.c_int, [_vp, ctypes.c_int64, ctypes.POINTER(KMeansParams), ctypes.c_int64, _vp, _vp, _vp, ctypes.c_int64, _i64p]),
    "pq_rand_perm": (None, [ctypes.c_int64, ctypes.c_int64, _vp]),
    "pq_kmeans_set_centroids": (ctypes.c_int, [_vp, ctypes.c_int64, _vp, ctypes.c_int]),
    "pq_kmeans_partial_device": (ctypes.c_int, [_vp, ctypes.c_int64, ctypes.c_int64, _vp, _vp, _vp, ctypes.POINTER(ctypes.c_double)]),
    "pq_kmeans_finish_device": (ctypes.c_int, [_vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, _vp, _vp, _vp, ctypes.POINTER(ctypes.c_int)]),
    "pq_plan_describe": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, _i64p, ctypes.c_int]),
    "pq_plan_describe_large_k": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, _i64p, ctypes.c_int]),
    "pq_last_error": (ctypes.c_char_p, []),
    "pq_version": (ctypes.c_char_p, []),
}


def library_path() -> str:
    return os.environ.get("PROQA_B200_LIB", os.path.join(_HERE, "libproqa_b200.so"))


def lib():
    """Load (once) and return the C-ABI library.  Raises if it has not been built."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise RuntimeError(
                f"proqa_b200: native library not found at {path}. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` or `make -C proqa_b200/csrc` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        L = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def last_error() -> str:
    return lib().pq_last_error().decode("utf-8", "replace")


def version() -> str:
    return lib().pq_version().decode()


def check(rc: int, what: str):
    """Map C-ABI status codes to the exceptions FAISS's SWIG layer raises (RuntimeError) / numpy-ish ValueError."""
    if rc == PQ_OK:
        return
    msg = f"proqa_b200: {what} failed ({rc}): {last_error()}"
    if rc == -1:
        raise ValueError(msg)
    if rc == -4:
        raise MemoryError(msg)
    raise RuntimeError(msg)
