"""faiss.Clustering on the B200 engine (host-side mirror of the reference's operator interface).

The reference (retrieval/group_paras.py:40-47) does

    clus = faiss.Clustering(d, ncentroids)
    clus.verbose = True; clus.niter = niter; clus.max_points_per_centroid = max_points_per_centroid
    clus.train(x, index)                                   # index = IndexFlatL2(d) or IndexFlatIP(d)
    centroids = faiss.vector_float_to_array(clus.centroids).reshape(ncentroids, d)

Same attribute names, defaults and messages as FAISS 1.6.3's Clustering/ClusteringParameters; the work happens in
pq_kmeans_train (proqa_b200/csrc/pq_kmeans.cu) behind the C ABI.  No CPU fallback.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .index import IndexFlat


class ClusteringParameters:
    def __init__(self):
        self.niter = 25
        self.nredo = 1
        self.verbose = False
        self.spherical = False
        self.update_index = False
        self.frozen_centroids = False
        self.min_points_per_centroid = 39
        self.max_points_per_centroid = 256
        self.seed = 1234


class Clustering(ClusteringParameters):
    def __init__(self, d, k, cp=None):
        super().__init__()
        if cp is not None:
            self.__dict__.update(cp.__dict__)
        self.d, self.k = int(d), int(k)
        self.centroids = np.zeros(0, dtype=np.float32)   # faiss: std::vector<float>, k*d after training
        self.obj = np.zeros(0, dtype=np.float32)         # objective per iteration
        self._own_result = None                          # the centroids array the last train() of this object produced

    def train(self, x, index):
        from .multi import MultiGpuIndexFlat
        multi = index if isinstance(index, MultiGpuIndexFlat) else None
        assert multi is not None or isinstance(index, IndexFlat), "proqa_b200.Clustering trains on the engine's IndexFlatL2 / IndexFlatIP"
        handle = multi._first_shard_handle() if multi is not None else index._h
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2, "train expects a 2-D array"
        n, d = x.shape
        assert d == self.d, f"dimension mismatch: got {d}, clustering has {self.d}"
        # FAISS treats centroids present before train() as input centroids.  ProQA never supplies any; what a previous train()
        # of this object left there is simply replaced (training starts afresh).
        if self.frozen_centroids or (len(self.centroids) and self.centroids is not self._own_result):
            raise NotImplementedError("proqa_b200.Clustering: input / frozen centroids are not supported (unused by ProQA)")
        prm = _lib.KMeansParams()
        _lib.lib().pq_kmeans_default_params(ctypes.byref(prm))
        prm.niter, prm.nredo = int(self.niter), int(self.nredo)
        prm.verbose, prm.spherical = int(bool(self.verbose)), int(bool(self.spherical))
        prm.min_points_per_centroid, prm.max_points_per_centroid = int(self.min_points_per_centroid), int(self.max_points_per_centroid)
        prm.seed = int(self.seed)
        cent = np.empty(self.k * self.d, dtype=np.float32)
        obj = np.zeros(max(1, prm.niter), dtype=np.float32)
        n_obj = ctypes.c_int64(0)
        if multi is not None:
            multi.reset()          # (the shard on the first device trains; the replicas are refilled below)
        rc = _lib.lib().pq_kmeans_train(handle, self.k, ctypes.byref(prm), n, x.ctypes.data, cent.ctypes.data, obj.ctypes.data, len(obj),
                                        ctypes.byref(n_obj))
        if rc == -4:
            raise MemoryError(f"proqa_b200: Clustering.train failed ({rc}): {_lib.last_error()}")
        if rc != 0:   # FAISS_THROW_IF_NOT_* (too few points, NaN input) reaches Python as RuntimeError through the SWIG layer
            raise RuntimeError(f"proqa_b200: Clustering.train failed ({rc}): {_lib.last_error()}")
        self.centroids = self._own_result = cent
        self.obj = obj[:n_obj.value].copy()
        if multi is not None:      # FAISS leaves the index holding the final centroids: on every GPU here
            multi.reset()
            multi.add(cent.reshape(self.k, self.d))


def vector_float_to_array(v):
    """faiss.vector_float_to_array (group_paras.py:46): a numpy copy of a std::vector<float>."""
    return np.array(v, dtype=np.float32, copy=True)


def vector_to_array(v):
    return np.array(v, copy=True)
