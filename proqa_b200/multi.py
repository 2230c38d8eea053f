"""One process, several GPUs — the multi-GPU index behind the FAISS-style API (C side: proqa_b200/csrc/pq_multi.cu).

The reference drives its index from one Python process (retrieval/eval_retrieval.py:102-104, retrieval/group_paras.py:36-51).
``MultiGpuIndexFlat`` has the surface of ``IndexFlat`` and spreads the work over the GPUs of the box with one host thread per
device inside the native library (no torch, no NCCL, no second process):

* a first ``add`` of more than 2**20 rows shards the rows contiguously over the devices; ``search`` replicates the queries,
  every shard searches its rows while the shards exchange thresholds through each other's HBM over NVLink, the per-shard
  lists are gathered on the first device and merged there (north_star (4));
* a smaller first ``add`` (k-means centroids, group_paras.py:49-51) is replicated and ``search`` splits the QUERIES instead.

The faiss shim returns this class from ``faiss.IndexFlatIP(d)`` / ``IndexFlatL2(d)`` when ``PROQA_B200_DEVICES`` names more
than one device (e.g. ``PROQA_B200_DEVICES=0,1,2,3,4,5,6,7`` or ``all``), so the unmodified scripts use the whole box.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import _lib
from .index import METRIC_INNER_PRODUCT, METRIC_L2, _LAST_STATS


def devices_from_env():
    """PROQA_B200_DEVICES: comma-separated ordinals, or 'all'.  None when unset / a single device."""
    s = os.environ.get("PROQA_B200_DEVICES", "").strip()
    if not s:
        return None
    if s.lower() == "all":
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis is not None:
            n = len([v for v in vis.split(",") if v.strip()])
        else:   # counting devices must not initialise CUDA in this process (the reference forks after importing faiss)
            n = len([f for f in os.listdir("/proc/driver/nvidia/gpus")]) if os.path.isdir("/proc/driver/nvidia/gpus") else 1
        devs = list(range(n))
    else:
        devs = [int(v) for v in s.split(",") if v.strip()]
    return devs if len(devs) > 1 else None


class MultiGpuIndexFlat:
    def __init__(self, d, metric=METRIC_INNER_PRODUCT, devices=(0,)):
        self.d, self.metric_type, self.is_trained, self.verbose = int(d), int(metric), True, False
        self.devices = [int(v) for v in devices]
        arr = (ctypes.c_int * len(self.devices))(*self.devices)
        self._h = ctypes.c_void_p()
        _lib.check(_lib.lib().pq_multi_create(self.d, self.metric_type, len(self.devices), arr, ctypes.byref(self._h)), "MultiGpuIndexFlat()")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().pq_multi_free(h)
            except Exception:
                pass

    @property
    def ntotal(self):
        return int(_lib.lib().pq_multi_ntotal(self._h))

    @property
    def layout(self):
        return {0: "empty", 1: "rows sharded", 2: "rows replicated, queries split"}[int(_lib.lib().pq_multi_mode(self._h))]

    def train(self, x):
        return None

    def add(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2, "add expects a 2-D array"
        n, d = x.shape
        assert d == self.d, f"dimension mismatch: got {d}, index has {self.d}"
        _lib.check(_lib.lib().pq_multi_add(self._h, n, x.ctypes.data), "add")

    def search(self, x, k):
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2, "search expects a 2-D array"
        n, d = x.shape
        assert d == self.d, f"dimension mismatch: got {d}, index has {self.d}"
        k = int(k)
        assert k > 0, "k must be positive"
        D = np.empty((n, k), dtype=np.float32)
        I = np.empty((n, k), dtype=np.int64)
        _lib.check(_lib.lib().pq_multi_search(self._h, n, x.ctypes.data, k, D.ctypes.data, I.ctypes.data), "search")
        buf = (ctypes.c_int64 * 10)()
        if _lib.lib().pq_multi_last_stats(self._h, buf, 10) == 0:
            _LAST_STATS[:] = list(buf)
        self.last_stats = list(buf)
        return D, I

    def reset(self):
        _lib.check(_lib.lib().pq_multi_reset(self._h), "reset")

    def _first_shard_handle(self):
        """pq_index* of the shard on the first device — where faiss.Clustering.train(x, index) runs its iterations."""
        return ctypes.c_void_p(_lib.lib().pq_multi_first_shard(self._h))


def make_index(d, metric):
    """What the faiss shim's IndexFlatIP / IndexFlatL2 construct: one GPU, or every GPU named by PROQA_B200_DEVICES."""
    devs = devices_from_env()
    if devs:
        return MultiGpuIndexFlat(d, metric, devs)
    from .index import IndexFlat
    return IndexFlat(d, metric)


__all__ = ["MultiGpuIndexFlat", "devices_from_env", "make_index", "METRIC_INNER_PRODUCT", "METRIC_L2"]
