"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU layer (shard bounds, id bases, all_gather layout,
merge order).  The per-rank searcher and the merge are test doubles built on the oracle — the product has no CPU path;
what is exercised here is proqa_b200/sharded.py itself."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import data


class _OracleLocal:
    """Stands in for the CUDA IndexFlat on one rank (same surface ShardedIndexFlat uses)."""

    def __init__(self, metric):
        from oracle import oracle
        self._o, self.metric, self.base, self.xb = oracle, metric, 0, np.zeros((0, 128), np.float32)

    def set_id_base(self, b):
        self.base = int(b)

    def add(self, x):
        self.xb = np.concatenate([self.xb, x])

    def reset(self):
        self.xb = np.zeros((0, 128), np.float32)

    def search(self, xq, k):
        return self._o.engine_spec(xq, self.xb, k, self.metric, id_base=self.base)


def _merge_double(D_all, I_all, k, metric):
    """What pq_merge_shard_results computes: best-first, ties -> lower global id, -1 padding last."""
    G, nq, _ = D_all.shape
    D = D_all.permute(1, 0, 2).reshape(nq, -1).numpy()
    I = I_all.permute(1, 0, 2).reshape(nq, -1).numpy()
    Do = np.empty((nq, k), np.float32)
    Io = np.empty((nq, k), np.int64)
    for q in range(nq):
        valid = I[q] >= 0
        key = -D[q] if metric == 0 else D[q]
        order = np.lexsort((I[q], key, ~valid))[:k]
        Do[q], Io[q] = D[q, order], I[q, order]
    return torch.from_numpy(Do), torch.from_numpy(Io)


def _merge_kernel_emulated(so_path):
    """The engine's own merge kernel (pq_merge_di_kernel + its launcher, what pq_merge_shard_results runs) executed under the
    SIMT emulator of tests/simt — so that the gloo path is checked with the real merge logic, not only with a double."""
    from tests.simt import harness
    lib = harness.load_select_emu(so_path)

    def merge(D_all, I_all, k, metric):
        D, I = harness.merge_di(lib, D_all.numpy(), I_all.numpy(), k, metric)
        return torch.from_numpy(D), torch.from_numpy(I)
    return merge


def _worker(rank, world, port, metric, n, nq, k, out, row_shards=None, merge_so=None):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from proqa_b200.sharded import ShardedIndexFlat, shard_bounds
        xb, xq = data.corpus(n), data.queries(nq)
        merge_fn = _merge_kernel_emulated(merge_so) if merge_so else _merge_double
        sh = ShardedIndexFlat(128, metric, local_factory=lambda: _OracleLocal(metric), merge_fn=merge_fn, row_shards=row_shards)
        sh.add(xb)
        R = row_shards or world
        lo, hi = shard_bounds(n, R, rank % R)
        assert sh.local.base == lo and len(sh.local.xb) == hi - lo and sh.ntotal == n
        D, I = sh.search(xq, k)
        np.savez(out + f".{rank}.npz", D=D, I=I)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("metric,n,nq,k", [(0, 3001, 7, 20), (1, 2000, 5, 8), (0, 30, 3, 40)])
def test_sharded_search_equals_single_index(tmp_path, metric, n, nq, k):
    from oracle import oracle
    out = str(tmp_path / "res")
    mp.spawn(_worker, args=(2, _free_port(), metric, n, nq, k, out), nprocs=2, join=True)
    Dr, Ir = oracle.engine_spec(data.queries(nq), data.corpus(n), k, metric)
    for rank in range(2):
        z = np.load(out + f".{rank}.npz")
        np.testing.assert_array_equal(z["I"], Ir)
        np.testing.assert_array_equal(z["D"].view(np.uint32), Dr.view(np.uint32))


@pytest.mark.parametrize("world,row_shards,nq", [(2, 1, 7), (2, 1, 8), (4, 2, 9), (4, 1, 5)])
def test_query_groups_times_row_shards_equals_single_index(tmp_path, world, row_shards, nq):
    """2-D layout: W ranks = Q query groups x R row shards (R = 1 is pure query sharding, ragged slices included)."""
    from oracle import oracle
    out = str(tmp_path / "res")
    n, k = 1500, 12
    mp.spawn(_worker, args=(world, _free_port(), 0, n, nq, k, out, row_shards), nprocs=world, join=True)
    Dr, Ir = oracle.engine_spec(data.queries(nq), data.corpus(n), k, 0)
    for rank in range(world):
        z = np.load(out + f".{rank}.npz")
        np.testing.assert_array_equal(z["I"], Ir)
        np.testing.assert_array_equal(z["D"].view(np.uint32), Dr.view(np.uint32))


@pytest.mark.parametrize("world,row_shards,metric,n,nq,k", [(2, None, 1, 900, 4, 50), (4, 2, 0, 1500, 9, 12)])
def test_sharded_search_through_the_engines_merge_kernel(tmp_path, world, row_shards, metric, n, nq, k):
    """Same as above with merge_fn = the engine's pq_merge_di_kernel under the SIMT emulator (built once, loaded by every rank)."""
    from oracle import oracle
    from tests.simt import harness
    so = harness.build_select_emu(tmp_path).path
    out = str(tmp_path / "res")
    mp.spawn(_worker, args=(world, _free_port(), metric, n, nq, k, out, row_shards, so), nprocs=world, join=True)
    Dr, Ir = oracle.engine_spec(data.queries(nq), data.corpus(n), k, metric)
    for rank in range(world):
        z = np.load(out + f".{rank}.npz")
        np.testing.assert_array_equal(z["I"], Ir)
        np.testing.assert_array_equal(z["D"].view(np.uint32), Dr.view(np.uint32))


def test_auto_row_shards():
    from proqa_b200.sharded import auto_row_shards
    assert auto_row_shards(8, 21_000_000) == 1            # 16 GB per copy: every GPU can hold the corpus
    assert auto_row_shards(8, 100_000_000) == 2           # 77 GB: two shards
    assert auto_row_shards(8, 1_000_000_000) == 8
    assert auto_row_shards(1, 10**9) == 1


def test_shard_bounds_cover_and_are_contiguous():
    from proqa_b200.sharded import shard_bounds
    for n in (0, 1, 7, 8, 9, 21_000_000, 100_000_001):
        for w in (1, 2, 4, 8):
            prev = 0
            for r in range(w):
                lo, hi = shard_bounds(n, w, r)
                assert lo == prev and hi >= lo
                prev = hi
            assert prev == n
