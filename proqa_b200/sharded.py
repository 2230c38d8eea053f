"""Multi-GPU layer: one process per GPU (torch.distributed), the corpus row-sharded over the ranks.

The reference's search path has no collective at all (single host process); sharding the rows of
``IndexFlatIP`` (retrieval/eval_retrieval.py:102-104) adds exactly one exchange step:

    rank r owns the contiguous global rows [lo_r, hi_r)   (ids reported as local row + lo_r)
    queries are replicated (broadcast from rank 0 when asked)
    every rank searches its shard              -> local (D, I) [nq, k], best-first, global ids
    all_gather of the R lists (NCCL over NVLink: 12 B per entry, tiny next to the scan)
    merge kernel (pq_merge_shard_results)      -> final (D, I) on every rank

Layout.  ``row_shards = R`` (default: the world size W, i.e. pure row sharding) arranges the W ranks as Q = W / R *query
groups* of R *row shards*: rank r holds row shard ``r % R`` and answers query slice ``r // R`` of Q.  Per-query work that
does not shrink with the shard (threshold epochs, selection, rescoring) is then divided by Q; the price is Q copies of the
corpus across the box (21M x 128 is 16 GB per copy, a B200 has 180).  The exchange becomes: all_gather + merge inside a
row group (R ranks), then one all_gather of the finished slices across the Q groups.  R = 1 is pure query sharding (no
merge at all); R = W is the layout above.

``local_factory`` / ``merge_fn`` exist so that the host-side logic (shard bounds, id bases, gather layout) can be
exercised on CPU with the gloo backend and test doubles; the defaults are the CUDA engine and there is no CPU fallback in
the product.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import _lib
from .index import METRIC_INNER_PRODUCT, IndexFlat


def shard_bounds(n, world, rank):
    """Contiguous, ascending, near-equal ranges: part ``rank`` of ``world`` owns [lo, hi)."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def auto_row_shards(world, n_rows, bytes_per_row=772, budget_bytes=48e9):
    """Fewest row shards (a divisor of ``world``) whose shard — fp32 rows + bf16 copy + norms — stays under the budget."""
    for r in range(1, world + 1):
        if world % r == 0 and (n_rows / r) * bytes_per_row <= budget_bytes:
            return r
    return world


class ShardedIndexFlat:
    def __init__(self, d, metric=METRIC_INNER_PRODUCT, group=None, device=None, local_factory=None, merge_fn=None, row_shards=None):
        import torch.distributed as dist
        self._dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.R = int(row_shards) if row_shards else self.world
        assert self.world % self.R == 0, f"row_shards={self.R} must divide the world size {self.world}"
        self.Q = self.world // self.R
        self.rr, self.rq = self.rank % self.R, self.rank // self.R   # my row shard, my query group
        self.d, self.metric_type, self.is_trained = int(d), int(metric), True
        self.device = device
        self._local = (local_factory or (lambda: IndexFlat(d, metric, -1 if device is None else device)))()
        self._merge_fn = merge_fn
        self.ntotal = 0
        self._first_add = True
        # threshold exchange between the row shards of a group (peer mailboxes over NVLink, include/proqa_b200.h); on by default
        # for the CUDA engine, PROQA_B200_SHARE=0 leaves the shards independent
        self._share_on = (merge_fn is None and local_factory is None and os.environ.get("PROQA_B200_SHARE", "1") != "0")
        self._share_cap = 0
        self._share_peers = []
        self._seq = 0
        # the exchange step itself: lists stored straight into the peers' HBM, merge spread over the ranks (pq_xchg.cu);
        # PROQA_B200_XCHG=0 falls back to NCCL all-gather + merge kernel on every rank
        self._xchg_on = (merge_fn is None and local_factory is None and os.environ.get("PROQA_B200_XCHG", "1") != "0")
        self._xchg = None
        self._xchg_bytes = 0
        self._xchg_peers = []
        self._xseq = 0
        self._local_is_engine = local_factory is None
        self._stream_set = False
        # process groups: every rank creates every group, in the same order (new_group is collective)
        self.row_group = group
        self.col_group = None
        if self.world > 1 and self.Q > 1:
            base = list(range(self.world)) if group is None else dist.get_process_group_ranks(group)
            for q in range(self.Q):
                g = dist.new_group([base[q * self.R + i] for i in range(self.R)]) if self.R > 1 else None
                if q == self.rq:
                    self.row_group = g
            for r in range(self.R):
                g = dist.new_group([base[q * self.R + r] for q in range(self.Q)])
                if r == self.rr:
                    self.col_group = g

    # ---- build ----------------------------------------------------------------------------------
    def row_bounds(self, n):
        return shard_bounds(n, self.R, self.rr)

    def query_bounds(self, nq):
        return shard_bounds(nq, self.Q, self.rq)

    def add(self, x):
        """Every rank passes the same global array; each keeps the contiguous slice of its row shard."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d
        lo, hi = self.row_bounds(x.shape[0])
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
        self._share_close()
        self._xchg_close()

    def close(self):
        """Collective: unmap the peers' buffers everywhere, then free the own ones (an exported buffer must outlive its mappings)."""
        import torch
        if self.world > 1 and self._merge_fn is None and self._local_is_engine:
            torch.cuda.synchronize()
            self._dist.barrier(group=self.group)
        L = _lib.lib()
        for ptr in self._share_peers:
            L.pq_ipc_close(self._dev_index(), ctypes.c_void_p(ptr))
        self._share_peers = []
        for ptr in self._xchg_peers:
            L.pq_ipc_close(self._dev_index(), ctypes.c_void_p(ptr))
        self._xchg_peers = []
        if self.world > 1 and self._merge_fn is None and self._local_is_engine:
            self._dist.barrier(group=self.group)
        self._share_close()
        self._xchg_close()

    # ---- threshold exchange between row shards ----------------------------------------------------
    def _share_close(self):
        L = _lib.lib()
        if self._share_cap:
            L.pq_index_share_close(self._local._h)
            for ptr in self._share_peers:
                L.pq_ipc_close(self._dev_index(), ctypes.c_void_p(ptr))
        self._share_cap, self._share_peers = 0, []

    def _dev_index(self):
        import torch
        return torch.cuda.current_device() if self.device is None else int(self.device)

    def _share_setup(self, nq):
        """Collective over the row group: (re)allocate the mailboxes for up to nq queries per search, exchange their CUDA IPC
        handles, map the peers' mailboxes, agree on the error-bound scalars of the whole corpus."""
        import torch
        L = _lib.lib()
        dist, dev = self._dist, torch.device("cuda", self._dev_index())
        self._share_close()
        cap = 1024
        while cap < min(nq, 1 << 18):
            cap *= 2
        sc = (ctypes.c_float * 2)()
        _lib.check(L.pq_index_get_bound_scalars(self._local._h, sc), "get_bound_scalars")
        t = torch.tensor([sc[0], sc[1]], dtype=torch.float32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.row_group)
        mx = t.cpu().tolist()
        _lib.check(L.pq_index_set_bound_scalars(self._local._h, mx[0], mx[1]), "set_bound_scalars")
        box, nbytes = ctypes.c_void_p(), ctypes.c_int64()
        _lib.check(L.pq_index_share_alloc(self._local._h, self.R, self.rr, cap, ctypes.byref(box), ctypes.byref(nbytes)), "share_alloc")
        handle = (ctypes.c_ubyte * 64)()
        _lib.check(L.pq_ipc_export(box, handle), "ipc_export")
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
        every = torch.empty(self.R * 64, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(every, mine, group=self.row_group)
        every = every.cpu().numpy().reshape(self.R, 64)
        ptrs = (ctypes.c_void_p * self.R)()
        for r in range(self.R):
            if r == self.rr:
                ptrs[r] = box.value
                continue
            h = (ctypes.c_ubyte * 64)(*every[r].tolist())
            out = ctypes.c_void_p()
            _lib.check(L.pq_ipc_open(h, self._dev_index(), ctypes.byref(out)), "ipc_open")
            ptrs[r] = out.value
            self._share_peers.append(out.value)
        _lib.check(L.pq_index_share_connect(self._local._h, ptrs), "share_connect")
        dist.barrier(group=self.row_group)   # nobody searches before every mailbox is mapped everywhere
        self._share_cap = cap

    # ---- list exchange over peer memory ------------------------------------------------------------
    def _xchg_close(self):
        L = _lib.lib()
        for ptr in self._xchg_peers:
            L.pq_ipc_close(self._dev_index(), ctypes.c_void_p(ptr))
        if self._xchg is not None:
            L.pq_xchg_free(self._xchg)
        self._xchg, self._xchg_bytes, self._xchg_peers = None, 0, []

    def _xchg_setup(self, nq, k):
        """Collective over all ranks: (re)allocate the exchange buffers for [nq, k] results and map every rank's buffer here."""
        import torch
        L = _lib.lib()
        dist, dev = self._dist, torch.device("cuda", self._dev_index())
        torch.cuda.synchronize()
        dist.barrier(group=self.group)       # nobody is still inside an exchange that uses the old buffers
        self._xchg_close()
        need = int(L.pq_xchg_bytes_needed(nq, k, self.R, self.Q))
        need = max(need, 1 << 20) * 5 // 4   # head-room: a slightly larger batch later does not re-map everything
        h, base, nbytes = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int64()
        _lib.check(L.pq_xchg_create(self._dev_index(), self.world, self.rank, need, ctypes.byref(h), ctypes.byref(base), ctypes.byref(nbytes)),
                   "xchg_create")
        handle = (ctypes.c_ubyte * 64)()
        _lib.check(L.pq_ipc_export(base, handle), "ipc_export")
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
        every = torch.empty(self.world * 64, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(every, mine, group=self.group)
        every = every.cpu().numpy().reshape(self.world, 64)
        ptrs = (ctypes.c_void_p * self.world)()
        for r in range(self.world):
            if r == self.rank:
                ptrs[r] = base.value
                continue
            out = ctypes.c_void_p()
            _lib.check(L.pq_ipc_open((ctypes.c_ubyte * 64)(*every[r].tolist()), self._dev_index(), ctypes.byref(out)), "ipc_open")
            ptrs[r] = out.value
            self._xchg_peers.append(out.value)
        _lib.check(L.pq_xchg_connect(h, ptrs), "xchg_connect")
        dist.barrier(group=self.group)
        self._xchg, self._xchg_bytes = h, need

    def _exchange(self, Dl, Il, nq, k, D_out, I_out):
        """Local list of this rank's query slice -> the full [nq, k] result on every rank, on torch's current stream."""
        import torch
        L = _lib.lib()
        if int(L.pq_xchg_bytes_needed(nq, k, self.R, self.Q)) > self._xchg_bytes:
            self._xchg_setup(nq, k)
        self._xseq += 1
        rc = L.pq_xchg_run(self._xchg, self.metric_type, self.R, nq, k, ctypes.c_void_p(Dl.data_ptr() if Dl.numel() else 0),
                           ctypes.c_void_p(Il.data_ptr() if Il.numel() else 0), ctypes.c_void_p(D_out.data_ptr()), ctypes.c_void_p(I_out.data_ptr()),
                           self._xseq, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "xchg_run")

    def _share_begin(self, nq_local):
        """Before every search of this rank's query slice (same on all ranks of the row group)."""
        if not (self._share_on and self.R > 1 and self.world > 1):
            return
        if min(nq_local, 1 << 18) > self._share_cap:
            self._share_setup(nq_local)
        self._seq += 1
        _lib.check(_lib.lib().pq_index_share_begin(self._local._h, self._seq), "share_begin")

    @property
    def local(self):
        return self._local

    # ---- search ---------------------------------------------------------------------------------
    def search(self, xq, k):
        """Host API (numpy in, numpy out), same result on every rank."""
        import torch
        xq = np.ascontiguousarray(xq, dtype=np.float32)
        k = int(k)
        nq = xq.shape[0]
        qlo, qhi = self.query_bounds(nq)
        if self.world == 1:
            return self._local.search(xq, k)
        if self._merge_fn is None and self._local_is_engine:
            # queries to the device once, local search + exchange on torch's current stream, one read-back of the result
            dev = torch.device("cuda", self._dev_index())
            if not self._stream_set:
                self._local.set_stream(torch.cuda.current_stream().cuda_stream)
                self._stream_set = True
            q = torch.from_numpy(xq).to(dev, non_blocking=True)
            n_loc = qhi - qlo
            D_loc = torch.empty((max(n_loc, 1), k), dtype=torch.float32, device=dev)
            I_loc = torch.empty((max(n_loc, 1), k), dtype=torch.int64, device=dev)
            D_out = torch.empty((nq, k), dtype=torch.float32, device=dev)
            I_out = torch.empty((nq, k), dtype=torch.int64, device=dev)
            D_all = I_all = None
            if not self._xchg_on:
                D_all = torch.empty((self.world, nq, k), dtype=torch.float32, device=dev)
                I_all = torch.empty((self.world, nq, k), dtype=torch.int64, device=dev)
            self.search_device(q, k, D_loc, I_loc, D_all, I_all, D_out, I_out)
            Dh, Ih = D_out.cpu().numpy(), I_out.cpu().numpy()
            if self._xchg_on:
                _lib.check(_lib.lib().pq_xchg_check(self._xchg), "list exchange")
            return Dh, Ih
        self._share_begin(qhi - qlo)
        D, I = self._local.search(xq[qlo:qhi], k)
        on_gpu = self._merge_fn is None
        dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
        Dl, Il = torch.from_numpy(D).to(dev), torch.from_numpy(I).to(dev)
        if self.R > 1:
            Dl, Il = self.gather_merge(Dl, Il, k)
        if self.Q > 1:
            Dl, Il = self._gather_slices(Dl, Il, nq, k)
        return Dl.cpu().numpy(), Il.cpu().numpy()

    def search_device(self, q, k, D_local, I_local, D_all, I_all, D_out, I_out):
        """Device API on torch tensors (all preallocated for the full batch): local search of this rank's query slice,
        all_gather + merge kernel inside the row group, all_gather of the slices across query groups."""
        nq = q.shape[0]
        qlo, qhi = self.query_bounds(nq)
        n_loc = qhi - qlo
        per = (nq + self.Q - 1) // self.Q   # padded slice length (all_gather needs equal pieces)
        Dl, Il = D_local[:n_loc], I_local[:n_loc]
        self._share_begin(n_loc)
        if n_loc:
            self._local.search_device(q[qlo:qhi].data_ptr(), n_loc, k, Dl.data_ptr(), Il.data_ptr())
        if self.world == 1:
            D_out.copy_(D_local)
            I_out.copy_(I_local)
            return
        if self._xchg_on:
            self._exchange(Dl, Il, nq, k, D_out, I_out)
            return
        if self.R > 1 and n_loc:  # (every rank of a row group has the same slice, so an empty slice skips consistently)
            # rank-major concatenation along dim 0: [R*n_loc, k] is the layout every backend accepts for the output
            Da, Ia = D_all.view(-1)[: self.R * n_loc * k].view(self.R * n_loc, k), I_all.view(-1)[: self.R * n_loc * k].view(self.R * n_loc, k)
            self._dist.all_gather_into_tensor(Da, Dl, group=self.row_group)
            self._dist.all_gather_into_tensor(Ia, Il, group=self.row_group)
            if self.Q == 1:
                self._merge_device(Da, Ia, n_loc, k, D_out, I_out)
                return
            self._merge_device(Da, Ia, n_loc, k, Dl, Il)   # slice result back into the local buffers
        if self.Q > 1:
            if per * self.Q == nq:
                self._dist.all_gather_into_tensor(D_out, D_local[:per], group=self.col_group)
                self._dist.all_gather_into_tensor(I_out, I_local[:per], group=self.col_group)
            else:  # ragged last slice: gather padded pieces, then trim
                Dp, Ip = self._gather_slices(D_local[:per], I_local[:per], nq, k, padded=True)
                D_out.copy_(Dp[:nq])
                I_out.copy_(Ip[:nq])

    def gather_merge(self, Dl, Il, k):
        """all_gather of the R per-shard lists of this row group + merge."""
        import torch
        nq = Dl.shape[0]
        D_all = torch.empty((self.R,) + tuple(Dl.shape), dtype=Dl.dtype, device=Dl.device)
        I_all = torch.empty((self.R,) + tuple(Il.shape), dtype=Il.dtype, device=Il.device)
        self._dist.all_gather_into_tensor(D_all.view(self.R * nq, k), Dl.contiguous(), group=self.row_group)
        self._dist.all_gather_into_tensor(I_all.view(self.R * nq, k), Il.contiguous(), group=self.row_group)
        if self._merge_fn is not None:
            return self._merge_fn(D_all, I_all, k, self.metric_type)
        D_out, I_out = torch.empty_like(Dl), torch.empty_like(Il)
        self._merge_device(D_all.view(self.R * nq, k), I_all.view(self.R * nq, k), nq, k, D_out, I_out)
        return D_out, I_out

    def _gather_slices(self, Dl, Il, nq, k, padded=False):
        """Finished query slices of the Q groups -> the full [nq, k] result on every rank."""
        import torch
        per = (nq + self.Q - 1) // self.Q
        if not padded:
            Dp = torch.zeros((per, k), dtype=Dl.dtype, device=Dl.device)
            Ip = torch.full((per, k), -1, dtype=Il.dtype, device=Il.device)
            Dp[: Dl.shape[0]] = Dl
            Ip[: Il.shape[0]] = Il
        else:
            Dp, Ip = Dl, Il
        D_all = torch.empty((self.Q * per, k), dtype=Dl.dtype, device=Dl.device)
        I_all = torch.empty((self.Q * per, k), dtype=Il.dtype, device=Il.device)
        self._dist.all_gather_into_tensor(D_all, Dp.contiguous(), group=self.col_group)
        self._dist.all_gather_into_tensor(I_all, Ip.contiguous(), group=self.col_group)
        return (D_all, I_all) if padded else (D_all[:nq], I_all[:nq])

    def _merge_device(self, D_all, I_all, nq, k, D_out, I_out):
        """Merge kernel on torch's current stream: ordered after the all-gather, no host synchronisation."""
        import torch
        stream = torch.cuda.current_stream().cuda_stream
        rc = _lib.lib().pq_merge_shard_results_async(D_all.device.index, self.metric_type, self.R, nq, k,
                                                     ctypes.c_void_p(D_all.data_ptr()), ctypes.c_void_p(I_all.data_ptr()),
                                                     ctypes.c_void_p(D_out.data_ptr()), ctypes.c_void_p(I_out.data_ptr()),
                                                     ctypes.c_void_p(stream))
        _lib.check(rc, "merge_shard_results")
