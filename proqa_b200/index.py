"""FAISS-style flat indexes on the B200 engine (host-side mirror of the reference's operator interface).

Same names, argument meaning and error behaviour as the FAISS 1.6.3 Python API the reference calls:

    index = IndexFlatIP(d)            # eval_retrieval.py:102, group_paras.py:36, trec_process.py:74
    index = IndexFlatL2(d)            # group_paras.py:38
    index.add(xb)                     # eval_retrieval.py:103  — float32 [n, d]; copies; ids = insertion order
    D, I = index.search(xq, k)        # eval_retrieval.py:104  — new float32 [nq,k] / int64 [nq,k], best first,
                                      #                          -1 / -FLT_MAX padding when k > ntotal
    index.reset()                     # group_paras.py:49
    index.ntotal, index.d, index.is_trained, index.metric_type

FAISS's wrapper asserts shapes (AssertionError) and surfaces C++ failures as RuntimeError; so does this.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1

_LAST_STATS = [0] * 10


def last_search_stats():
    """Counters of the most recent search in this process (see pq_index_last_stats in include/proqa_b200.h)."""
    keys = ["tc_queries", "fp32_rerun_queries", "fp32_scan_launches", "tc_filter_launches", "select_launches", "kernel_launches",
            "device_us", "dominant_kernel_us", "epoch_second_attempts", "reserved"]
    return dict(zip(keys, _LAST_STATS))


class IndexFlat:
    def __init__(self, d, metric=METRIC_INNER_PRODUCT, device=-1):
        self.d = int(d)
        self.metric_type = int(metric)
        self.is_trained = True
        self.verbose = False
        self._h = ctypes.c_void_p()
        _lib.check(_lib.lib().pq_index_create(self.d, self.metric_type, int(device), ctypes.byref(self._h)), "IndexFlat()")

    # -- lifetime ---------------------------------------------------------------------------------
    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().pq_index_free(h)
            except Exception:
                pass

    # -- FAISS surface ----------------------------------------------------------------------------
    @property
    def ntotal(self):
        return int(_lib.lib().pq_index_ntotal(self._h))

    def train(self, x):
        """Flat indexes need no training (FAISS: is_trained is always True)."""
        return None

    def add(self, x):
        """float32 [n, d] as FAISS requires; a float16 array (embeddings saved by get_embed.py --fp16, before the
        .astype('float32') of eval_retrieval.py:100) is accepted too and widened on the device."""
        if getattr(x, "dtype", None) == np.float16:
            x = np.ascontiguousarray(x)
            assert x.ndim == 2, "add expects a 2-D array"
            n, d = x.shape
            assert d == self.d, f"dimension mismatch: got {d}, index has {self.d}"
            _lib.check(_lib.lib().pq_index_add_f16(self._h, n, x.ctypes.data), "add")
            return
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2, "add expects a 2-D array"
        n, d = x.shape
        assert d == self.d, f"dimension mismatch: got {d}, index has {self.d}"
        _lib.check(_lib.lib().pq_index_add(self._h, n, x.ctypes.data), "add")

    def search(self, x, k):
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2, "search expects a 2-D array"
        n, d = x.shape
        assert d == self.d, f"dimension mismatch: got {d}, index has {self.d}"
        k = int(k)
        assert k > 0, "k must be positive"
        D = np.empty((n, k), dtype=np.float32)
        I = np.empty((n, k), dtype=np.int64)
        _lib.check(_lib.lib().pq_index_search(self._h, n, x.ctypes.data, k, D.ctypes.data, I.ctypes.data), "search")
        self._pull_stats()
        return D, I

    def reset(self):
        _lib.check(_lib.lib().pq_index_reset(self._h), "reset")

    def add_npy(self, path, chunk_rows=1 << 20):
        """Append the rows of a .npy file (float32 or float16 [n, d], as written by retrieval/get_embed.py:139) without
        materialising it in host memory: the file is memory-mapped and streamed chunk by chunk through the pinned staging
        buffers; float16 files are widened on the device.  Equivalent to ``index.add(np.load(path).astype('float32'))``
        (eval_retrieval.py:100,103) minus the two full-size host copies."""
        x = np.load(path, mmap_mode="r")
        assert x.ndim == 2 and x.shape[1] == self.d, f"expected [n, {self.d}], file has shape {x.shape}"
        assert x.dtype in (np.float32, np.float16), f"unsupported dtype {x.dtype}"
        for a in range(0, x.shape[0], int(chunk_rows)):
            self.add(np.ascontiguousarray(x[a:a + int(chunk_rows)]))
        return x.shape[0]

    # -- engine extensions (not part of FAISS) ----------------------------------------------------
    def add_device(self, ptr, n):
        """Append n rows that already live in device memory (raw float32 device pointer, C-contiguous [n, d])."""
        _lib.check(_lib.lib().pq_index_add_device(self._h, int(n), ctypes.c_void_p(int(ptr))), "add_device")

    def search_device(self, q_ptr, nq, k, D_ptr, I_ptr):
        """Search with queries and outputs in device memory (raw pointers: float32 [nq,d], float32 [nq,k], int64 [nq,k])."""
        _lib.check(_lib.lib().pq_index_search_device(self._h, int(nq), ctypes.c_void_p(int(q_ptr)), int(k), ctypes.c_void_p(int(D_ptr)),
                                                     ctypes.c_void_p(int(I_ptr))), "search_device")
        self._pull_stats()

    def set_id_base(self, base):
        _lib.check(_lib.lib().pq_index_set_id_base(self._h, int(base)), "set_id_base")

    def set_tier(self, tier):
        """'auto' | 'fp32' | 'bf16'  (see enum pq_tier)."""
        t = {"auto": _lib.TIER_AUTO, "fp32": _lib.TIER_FP32, "bf16": _lib.TIER_BF16}[tier] if isinstance(tier, str) else int(tier)
        _lib.check(_lib.lib().pq_index_set_tier(self._h, t), "set_tier")

    def set_stream(self, cuda_stream):
        """Run on the caller's CUDA stream (an int handle, e.g. torch.cuda.current_stream().cuda_stream); None = own stream."""
        if cuda_stream is None:
            _lib.check(_lib.lib().pq_index_set_stream(self._h, None, 0), "set_stream")
        else:
            _lib.check(_lib.lib().pq_index_set_stream(self._h, ctypes.c_void_p(int(cuda_stream)), 1), "set_stream")

    def set_profile(self, on=True):
        """Time the dominant kernel with CUDA events on the launching stream (last_stats[7], microseconds)."""
        _lib.check(_lib.lib().pq_index_set_profile(self._h, 1 if on else 0), "set_profile")

    def _pull_stats(self):
        buf = (ctypes.c_int64 * 16)()
        if _lib.lib().pq_index_last_stats(self._h, buf, 16) == 0:
            _LAST_STATS[:] = list(buf)[:10]
        self.last_stats = list(buf)


class IndexFlatIP(IndexFlat):
    def __init__(self, d, device=-1):
        super().__init__(d, METRIC_INNER_PRODUCT, device)


class IndexFlatL2(IndexFlat):
    def __init__(self, d, device=-1):
        super().__init__(d, METRIC_L2, device)
