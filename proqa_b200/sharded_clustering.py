"""faiss.Clustering with the training POINTS sharded over several GPUs (one process per GPU, torch.distributed).

The reference trains k-means in one process (retrieval/group_paras.py:40-45: ``clus.train(x, index)``, 250 iterations over
up to ncentroids * max_points_per_centroid = 10M points).  The assignment ``index.search(x, 1)`` is independent per point,
so the points shard with no exchange; the centroid update needs exactly one all-reduce per iteration: the [k,128] fp32
sums, the [k] counts and the objective.

    every rank:  same x (e.g. the same memory-mapped .npy), same seed
      subsample + initial centroids exactly as FAISS does (rand_perm(seed), rand_perm(seed + 1)) -> identical on every rank
      rank r keeps the contiguous slice [lo_r, hi_r) of the (sub-sampled) training set on its GPU
      per iteration:  partial (assign local points, local sums/counts)  ->  all_reduce(SUM)  ->  finish (divide, split void
                      clusters, renormalise, index.reset(); index.add(centroids)) — identical totals, identical centroids

Differences from the single-GPU ``Clustering`` (which restates FAISS bit for bit on well-separated data): a centroid's sum
is formed per shard in index order and the shard sums are then added by NCCL, so the last bits of a centroid can differ from
the sequential sum; assignments can only change where two centroids are within that rounding of each other.

``backend`` exists so that the host logic (sharding, all-reduce, iteration order) can be exercised on CPU with gloo and a
test double; the default backend is the CUDA engine (pq_kmeans_* in include/proqa_b200.h) and there is no CPU fallback.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .clustering import ClusteringParameters
from .sharded import shard_bounds


class _EngineBackend:
    """The three device steps behind the C ABI; tensors live on this rank's GPU."""

    def __init__(self, index, k):
        import torch
        self.torch, self.index, self.k = torch, index, int(k)
        self.x = None
        dev = torch.device("cuda", torch.cuda.current_device())
        self.sums = torch.zeros((self.k, 128), dtype=torch.float32, device=dev)
        self.counts = torch.zeros((self.k,), dtype=torch.int32, device=dev)

    def rand_perm(self, n, seed):
        out = np.empty(n, dtype=np.int32)
        _lib.lib().pq_rand_perm(n, seed, ctypes.c_void_p(out.ctypes.data))
        return out

    def set_points(self, x_local):
        self.x = self.torch.from_numpy(np.ascontiguousarray(x_local, dtype=np.float32)).to(self.sums.device)

    def set_centroids(self, cent, spherical):
        cent = np.ascontiguousarray(cent, dtype=np.float32)
        _lib.check(_lib.lib().pq_kmeans_set_centroids(self.index._h, self.k, ctypes.c_void_p(cent.ctypes.data), int(spherical)), "kmeans_set_centroids")

    def partial(self):
        obj = ctypes.c_double(0.0)
        self.torch.cuda.current_stream().synchronize()      # the engine works on its own stream
        rc = _lib.lib().pq_kmeans_partial_device(self.index._h, self.k, self.x.shape[0], ctypes.c_void_p(self.x.data_ptr()),
                                                 ctypes.c_void_p(self.sums.data_ptr()), ctypes.c_void_p(self.counts.data_ptr()), ctypes.byref(obj))
        _lib.check(rc, "kmeans_partial_device")
        return self.sums, self.counts, obj.value

    def finish(self, sums, counts, n_total, spherical):
        cent = np.empty((self.k, 128), dtype=np.float32)
        nsplit = ctypes.c_int(0)
        self.torch.cuda.current_stream().synchronize()      # the all-reduce must have landed
        rc = _lib.lib().pq_kmeans_finish_device(self.index._h, self.k, n_total, int(spherical), ctypes.c_void_p(sums.data_ptr()),
                                                ctypes.c_void_p(counts.data_ptr()), ctypes.c_void_p(cent.ctypes.data), ctypes.byref(nsplit))
        _lib.check(rc, "kmeans_finish_device")
        return cent, nsplit.value


class ShardedClustering(ClusteringParameters):
    """Same attributes as ``Clustering``; ``train(x, index)`` is collective: every rank calls it with the same ``x``."""

    def __init__(self, d, k, cp=None, group=None, backend_factory=None):
        super().__init__()
        if cp is not None:
            self.__dict__.update(cp.__dict__)
        import torch.distributed as dist
        self._dist, self.group = dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.d, self.k = int(d), int(k)
        self._backend_factory = backend_factory
        self.centroids = np.zeros(0, dtype=np.float32)
        self.obj = np.zeros(0, dtype=np.float32)

    def train(self, x, index):
        import torch
        assert x.ndim == 2 and x.shape[1] == self.d == 128, f"dimension mismatch: got {x.shape}, clustering has d={self.d}"
        if self.nredo != 1:
            raise NotImplementedError("ShardedClustering: nredo > 1 is not supported (ProQA leaves it at 1)")
        n, k = int(x.shape[0]), self.k
        if n < k:
            raise RuntimeError(f"Number of training points ({n}) should be at least as large as number of clusters ({k})")
        # Clustering::train refuses non-finite input before anything else (as the single-GPU driver does); every rank holds the
        # same x, so each checks its own stripe and the verdict is shared
        flo, fhi = shard_bounds(n, self.world, self.rank)
        bad = not bool(np.isfinite(np.asarray(x[flo:fhi], dtype=np.float32)).all())
        if self.world > 1:
            flags = [None] * self.world
            self._dist.all_gather_object(flags, bad, group=self.group)
            bad = any(flags)
        if bad:
            raise RuntimeError("input contains NaN's or Inf's")
        be = (self._backend_factory or _EngineBackend)(index, k)
        # ---- subsample + initial centroids: FAISS's own draws, identical on every rank -----------------------------------
        max_train = k * int(self.max_points_per_centroid)
        if n > max_train:
            if self.verbose and self.rank == 0:
                print(f"Sampling a subset of {max_train} / {n} for training")
            sub = be.rand_perm(n, int(self.seed))[:max_train].astype(np.int64)
            nx = max_train
        else:
            sub, nx = None, n
        rows = (lambda idx: np.asarray(x[np.sort(idx)], dtype=np.float32)[np.argsort(np.argsort(idx))]) if sub is not None else None
        lo, hi = shard_bounds(nx, self.world, self.rank)
        x_local = rows(sub[lo:hi]) if sub is not None else np.asarray(x[lo:hi], dtype=np.float32)
        be.set_points(x_local)
        first = be.rand_perm(nx, int(self.seed) + 1)[:k].astype(np.int64)
        cent0 = rows(sub[first]) if sub is not None else np.stack([np.asarray(x[i], dtype=np.float32) for i in first])
        if nx == k:   # "Number of training points same as number of clusters, just copying" — no normalisation, no iterations
            if self.verbose and self.rank == 0:
                print(f"Number of training points ({nx}) same as number of clusters, just copying")
            pts = rows(sub) if sub is not None else np.asarray(x, dtype=np.float32)
            be.set_centroids(pts, False)
            self.centroids = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1)
            self.obj = np.zeros(0, dtype=np.float32)
            return
        be.set_centroids(cent0, self.spherical)
        if self.spherical:   # (niter = 0 must still hand back the normalised initial centroids, as Clustering::train does)
            cent0 = cent0 / np.maximum(np.sqrt((cent0.astype(np.float32) ** 2).sum(1, keepdims=True, dtype=np.float32)), np.float32(1e-30))
        if self.verbose and self.rank == 0:
            print(f"Clustering {nx} points in {self.d}D to {k} clusters, redo 1 times, {self.niter} iterations ({self.world} ranks)")
        # ---- iterations ------------------------------------------------------------------------------------------------------
        obj, cent = [], cent0
        for it in range(int(self.niter)):
            sums, counts, local_obj = be.partial()
            o = torch.tensor([local_obj], dtype=torch.float64, device=sums.device)
            if self.world > 1:
                self._dist.all_reduce(sums, group=self.group)
                self._dist.all_reduce(counts, group=self.group)
                self._dist.all_reduce(o, group=self.group)
            cent, nsplit = be.finish(sums, counts, nx, self.spherical)
            obj.append(np.float32(o.item()))
            if self.verbose and self.rank == 0:
                print(f"  Iteration {it}: objective={obj[-1]:g} nsplit={nsplit}")
        self.centroids = np.ascontiguousarray(cent, dtype=np.float32).reshape(-1)
        self.obj = np.array(obj, dtype=np.float32)
