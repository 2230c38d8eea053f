"""Multi-GPU layer: the corpus row-sharded over the ranks of a torch.distributed group (one process per GPU).

The reference's search path has no collective at all (single host process); sharding the
rows of ``IndexFlatIP`` (retrieval/eval_retrieval.py:102-104) adds exactly one exchange step:

    rank r owns the contiguous global rows [lo_r, hi_r)   (ids reported as local row + lo_r)
    queries are replicated (broadcast from rank 0 when asked)
    every rank searches its shard              -> local (D, I) [nq, k], best-first, global ids
    all_gather of the G lists (NCCL over NVLink: 12 B per entry, tiny next to the scan)
    merge kernel (pq_merge_shard_results)      -> final (D, I) on every rank

``local_factory`` / ``merge_fn`` exist so that the host-side logic (shard bounds, id bases, gather layout)
can be exercised on CPU with the gloo backend and test doubles; the defaults are the CUDA engine and there is
no CPU fallback in the product.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .index import METRIC_INNER_PRODUCT, IndexFlat


def shard_bounds(n, world, rank):
    """Contiguous, ascending, near-equal row ranges: rank r owns [lo, hi)."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


class ShardedIndexFlat:
    def __init__(self, d, metric=METRIC_INNER_PRODUCT, group=None, device=None, local_factory=None, merge_fn=None):
        import torch.distributed as dist
        self._dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.d, self.metric_type, self.is_trained = int(d), int(metric), True
        self.device = device
        self._local = (local_factory or (lambda: IndexFlat(d, metric, -1 if device is None else device)))()
        self._merge_fn = merge_fn
        self.ntotal = 0
        self._first_add = True

    # ---- build ----------------------------------------------------------------------------------
    def add(self, x):
        """Every rank passes the same global array; each keeps its contiguous slice."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d
        lo, hi = shard_bounds(x.shape[0], self.world, self.rank)
        self.add_shard(x[lo:hi], self.ntotal + lo, x.shape[0])

    def add_shard(self, x_local, id_base, n_global):
        """This rank's rows of a global append of n_global rows starting at global id ``id_base``."""
        if not self._first_add:
            raise NotImplementedError("ShardedIndexFlat: one contiguous range per rank (call add once, or reset first)")
        self._local.set_id_base(int(id_base))
        self._first_add = False
        if len(x_local):
            self._local.add(x_local)
        self.ntotal += int(n_global)

    def add_shard_device(self, ptr, n_local, id_base, n_global):
        if not self._first_add:
            raise NotImplementedError("ShardedIndexFlat: one contiguous range per rank")
        self._local.set_id_base(int(id_base))
        self._first_add = False
        self._local.add_device(ptr, n_local)
        self.ntotal += int(n_global)

    def reset(self):
        self._local.reset()
        self.ntotal = 0
        self._first_add = True

    @property
    def local(self):
        return self._local

    # ---- search ---------------------------------------------------------------------------------
    def search(self, xq, k):
        """Host API (numpy in, numpy out), same on every rank."""
        import torch
        xq = np.ascontiguousarray(xq, dtype=np.float32)
        k = int(k)
        D, I = self._local.search(xq, k)
        if self.world == 1:
            return D, I
        on_gpu = self._merge_fn is None
        dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
        Dl, Il = torch.from_numpy(D).to(dev), torch.from_numpy(I).to(dev)
        Dg, Ig = self.gather_merge(Dl, Il, k)
        return Dg.cpu().numpy(), Ig.cpu().numpy()

    def search_device(self, q, k, D_local, I_local, D_all, I_all, D_out, I_out):
        """Device API on torch tensors (all preallocated): local search, all_gather, merge kernel."""
        nq = q.shape[0]
        self._local.search_device(q.data_ptr(), nq, k, D_local.data_ptr(), I_local.data_ptr())
        if self.world == 1:
            D_out.copy_(D_local)
            I_out.copy_(I_local)
            return
        # rank-major concatenation along dim 0: [G*nq, k] is the layout every backend accepts for the output
        self._dist.all_gather_into_tensor(D_all.view(self.world * nq, k), D_local, group=self.group)
        self._dist.all_gather_into_tensor(I_all.view(self.world * nq, k), I_local, group=self.group)
        self._merge_device(D_all, I_all, nq, k, D_out, I_out)

    def gather_merge(self, Dl, Il, k):
        import torch
        nq = Dl.shape[0]
        D_all = torch.empty((self.world,) + tuple(Dl.shape), dtype=Dl.dtype, device=Dl.device)
        I_all = torch.empty((self.world,) + tuple(Il.shape), dtype=Il.dtype, device=Il.device)
        self._dist.all_gather_into_tensor(D_all.view(self.world * nq, k), Dl.contiguous(), group=self.group)
        self._dist.all_gather_into_tensor(I_all.view(self.world * nq, k), Il.contiguous(), group=self.group)
        if self._merge_fn is not None:
            return self._merge_fn(D_all, I_all, k, self.metric_type)
        D_out, I_out = torch.empty_like(Dl), torch.empty_like(Il)
        self._merge_device(D_all, I_all, nq, k, D_out, I_out)
        return D_out, I_out

    def _merge_device(self, D_all, I_all, nq, k, D_out, I_out):
        """Merge kernel on torch's current stream: ordered after the all-gather, no host synchronisation."""
        import torch
        stream = torch.cuda.current_stream().cuda_stream
        rc = _lib.lib().pq_merge_shard_results_async(D_all.device.index, self.metric_type, self.world, nq, k,
                                                     ctypes.c_void_p(D_all.data_ptr()), ctypes.c_void_p(I_all.data_ptr()),
                                                     ctypes.c_void_p(D_out.data_ptr()), ctypes.c_void_p(I_out.data_ptr()),
                                                     ctypes.c_void_p(stream))
        _lib.check(rc, "merge_shard_results")
