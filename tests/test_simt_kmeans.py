"""CPU: the k-means drivers of pq_kmeans.cu under the SIMT emulator (tests/simt) — the single-GPU driver against the FAISS
Clustering restatement, the staged steps (set_centroids / partial / finish: what multi-GPU training is made of) against the
single driver, and proqa_b200.sharded_clustering.ShardedClustering over gloo with those emulated steps as its backend.

Real: the drivers' own source text, FAISS's rand_perm / void-cluster split, and the kernels km_keys / km_centroid<mean|sum> /
km_divide / km_renorm / km_objective / km_gather_rows.  Stand-ins: CUB's radix sort (std::stable_sort) and the index
(brute force with the engine's defined score).  Test infrastructure only."""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle
from tests.simt import harness
from tests.test_kmeans_oracle import blobs

pytestmark = pytest.mark.timeout(600)   # an emulated kernel that never finishes must not hang the suite


@pytest.fixture(scope="module")
def km(tmp_path_factory):
    return harness.build_kmeans_emu(tmp_path_factory.mktemp("simt_kmeans"))


def emu_train(lib, x, k, niter, metric=1, spherical=0, max_pts=256, seed=1234):
    lib.emu_km_new(metric)
    cent = np.empty((k, 128), np.float32)
    obj = np.zeros(niter, np.float32)
    n_obj = ctypes.c_longlong(0)
    msg = lib.emu_km_train(x.ctypes.data, len(x), k, niter, spherical, max_pts, seed, cent.ctypes.data, obj.ctypes.data, niter, ctypes.byref(n_obj))
    assert msg is None, msg.decode()
    return cent, obj[:n_obj.value]


def emu_assign(lib, x):
    D = np.empty(len(x), np.float32)
    I = np.empty(len(x), np.int64)
    assert lib.emu_km_assign(x.ctypes.data, len(x), D.ctypes.data, I.ctypes.data) == 0
    return D, I


def oracle_train(x, k, niter, metric=1, spherical=False, max_pts=256):
    ix = oracle.FaissFlatOracle(128, metric)
    clus = oracle.FaissClusteringOracle(128, k)
    clus.niter, clus.spherical, clus.max_points_per_centroid = niter, spherical, max_pts
    clus.train(x, ix)
    return clus.centroids.reshape(k, 128), np.asarray(clus.obj, np.float32), ix


@pytest.mark.parametrize("metric,spherical", [(1, 0), (0, 1)])
def test_single_driver_matches_the_faiss_restatement(km, metric, spherical):
    x, _ = blobs(1500, 12, distinct_init=True)
    cent, obj = emu_train(km, x, 12, 3, metric=metric, spherical=spherical, max_pts=1000)
    cent_o, obj_o, ix = oracle_train(x, 12, 3, metric=metric, spherical=bool(spherical), max_pts=1000)
    if not spherical:   # unambiguous assignments + the same index-order fp32 sums and the same mean: the very same bits
        np.testing.assert_array_equal(cent.view(np.uint32), cent_o.view(np.uint32))
    np.testing.assert_allclose(cent, cent_o, rtol=2e-6, atol=2e-6)      # (spherical: the norm is summed in the engine's order)
    np.testing.assert_allclose(obj, obj_o, rtol=5e-4)                   # FAISS sums |x|^2+|y|^2-2xy in fp32: ~1e-4 absolute noise per distance
    _, I = emu_assign(km, x)
    np.testing.assert_array_equal(I, ix.search(x, 1)[1][:, 0])


def test_single_driver_subsamples_like_faiss(km):
    x, _ = blobs(3000, 4)
    cent, _ = emu_train(km, x, 4, 3, max_pts=100)                         # 400 of 3000 points
    cent_o, _, _ = oracle_train(x, 4, 3, max_pts=100)
    np.testing.assert_allclose(cent, cent_o, rtol=2e-6, atol=2e-6)


def test_void_clusters_are_split_like_faiss(km):
    from tests import data
    x = np.repeat(data.corpus(3, seed=9), 200, axis=0)                    # 3 distinct points, k = 8: most clusters start empty
    cent, obj = emu_train(km, x, 8, 3)
    cent_o, obj_o, _ = oracle_train(x, 8, 3)
    # duplicated points make duplicated centroids: which of two identical centroids wins a point is decided by rounding
    # (sgemm there, the engine's chain here), so only tie-independent facts are compared
    assert np.isfinite(cent).all()
    assert obj.max() < 1.0 and np.max(obj_o) < 1.0                        # every point coincides with a centroid throughout
    assert len(np.unique(np.round(cent, 2), axis=0)) >= 3


def staged_train(lib, shards, k, niter, metric, spherical, cent0, n_total):
    """The multi-GPU iteration on one process: partial per shard, sums added shard by shard (what the all-reduce does)."""
    lib.emu_km_new(metric)
    msg = lib.emu_km_set_centroids(k, cent0.ctypes.data, spherical)
    assert msg is None, msg
    objs = []
    cent = np.empty((k, 128), np.float32)
    for _ in range(niter):
        sums = np.zeros((k, 128), np.float32)
        counts = np.zeros(k, np.int32)
        total = 0.0
        for xs in shards:
            s = np.full((k, 128), np.nan, np.float32)
            c = np.full(k, -1, np.int32)
            o = ctypes.c_double(0)
            msg = lib.emu_km_partial(k, len(xs), xs.ctypes.data if len(xs) else None, s.ctypes.data, c.ctypes.data, ctypes.byref(o))
            assert msg is None, msg
            sums += s
            counts += c
            total += o.value
        nsplit = ctypes.c_int(0)
        msg = lib.emu_km_finish(k, n_total, spherical, sums.ctypes.data, counts.ctypes.data, cent.ctypes.data, ctypes.byref(nsplit))
        assert msg is None, msg
        objs.append(total)
    return cent.copy(), np.array(objs, np.float32)


@pytest.mark.parametrize("metric,spherical,n_shards", [(1, 0, 1), (1, 0, 3), (0, 1, 2)])
def test_staged_steps_match_the_single_driver(km, metric, spherical, n_shards):
    n, k, niter = 1500, 12, 3
    x, _ = blobs(n, k, distinct_init=True)
    cent_single, obj_single = emu_train(km, x, k, niter, metric=metric, spherical=spherical, max_pts=1000)
    first = oracle.rand_perm(n, 1235)[:k]
    bounds = [n * i // n_shards for i in range(n_shards + 1)]
    shards = [np.ascontiguousarray(x[a:b]) for a, b in zip(bounds, bounds[1:])]
    if n_shards == 3:
        shards.append(np.zeros((0, 128), np.float32))                     # a rank with no points at all
    cent, obj = staged_train(km, shards, k, niter, metric, spherical, np.ascontiguousarray(x[first]), n)
    if n_shards == 1:
        np.testing.assert_array_equal(cent, cent_single)                  # one shard: the very same sums
    else:
        np.testing.assert_allclose(cent, cent_single, rtol=3e-6, atol=3e-6)   # shard sums added in a different order
    np.testing.assert_allclose(obj, obj_single, rtol=5e-4)      # (distances are differences of O(2000) terms: last-bit centroid changes show)


# ---- ShardedClustering over gloo, its backend = the emulated engine steps -------------------------------------------------
class _EmuBackend:
    def __init__(self, lib, index, k):
        self.lib, self.k, self.metric = lib, k, index.metric_type
        lib.emu_km_new(self.metric)
        self.sums = torch.zeros((k, 128), dtype=torch.float32)
        self.counts = torch.zeros((k,), dtype=torch.int32)

    def rand_perm(self, n, seed):
        out = np.empty(n, np.int32)
        self.lib.emu_rand_perm(n, seed, out.ctypes.data)
        return out

    def set_points(self, x):
        self.x = np.ascontiguousarray(x, np.float32)

    def set_centroids(self, cent, spherical):
        cent = np.ascontiguousarray(cent, np.float32)
        assert self.lib.emu_km_set_centroids(self.k, cent.ctypes.data, int(spherical)) is None

    def partial(self):
        o = ctypes.c_double(0)
        assert self.lib.emu_km_partial(self.k, len(self.x), self.x.ctypes.data if len(self.x) else None, self.sums.data_ptr(), self.counts.data_ptr(),
                                       ctypes.byref(o)) is None
        return self.sums, self.counts, o.value

    def finish(self, sums, counts, n_total, spherical):
        cent = np.empty((self.k, 128), np.float32)
        ns = ctypes.c_int(0)
        assert self.lib.emu_km_finish(self.k, n_total, int(spherical), sums.data_ptr(), counts.data_ptr(), cent.ctypes.data, ctypes.byref(ns)) is None
        return cent, ns.value


class _IndexStub:
    def __init__(self, metric):
        self.metric_type = metric


def _worker(rank, world, port, so, metric, spherical, n, k, niter, max_pts, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from proqa_b200.sharded_clustering import ShardedClustering
        lib = harness.load_kmeans_emu(so)
        x, _ = blobs(n, k, distinct_init=(max_pts * k >= n))
        clus = ShardedClustering(128, k, backend_factory=lambda index, kk: _EmuBackend(lib, index, kk))
        clus.niter, clus.spherical, clus.max_points_per_centroid = niter, bool(spherical), max_pts
        clus.train(x, _IndexStub(metric))
        _, I = emu_assign(lib, x)
        np.savez(out + f".{rank}.npz", centroids=clus.centroids, obj=clus.obj, I=I)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,metric,spherical,n,k,niter,max_pts", [(2, 1, 0, 1500, 12, 3, 1000), (3, 0, 1, 3000, 4, 3, 100)])
def test_sharded_clustering_over_gloo_matches_single_process(tmp_path, km, world, metric, spherical, n, k, niter, max_pts):
    out = str(tmp_path / "res")
    mp.spawn(_worker, args=(world, _free_port(), km.path, metric, spherical, n, k, niter, max_pts, out), nprocs=world, join=True)
    x, _ = blobs(n, k, distinct_init=(max_pts * k >= n))
    cent_o, obj_o, ix = oracle_train(x, k, niter, metric=metric, spherical=bool(spherical), max_pts=max_pts)
    res = [np.load(out + f".{r}.npz") for r in range(world)]
    for z in res[1:]:                                                      # identical totals -> identical centroids on every rank
        np.testing.assert_array_equal(z["centroids"], res[0]["centroids"])
        np.testing.assert_array_equal(z["obj"], res[0]["obj"])
    np.testing.assert_allclose(res[0]["centroids"].reshape(k, 128), cent_o, rtol=5e-6, atol=5e-6)
    np.testing.assert_allclose(res[0]["obj"], obj_o, rtol=1e-3)            # (sgemm-expanded distances there, the engine's score here)
    np.testing.assert_array_equal(res[0]["I"], ix.search(x, 1)[1][:, 0])


def test_group_paras_flow_on_the_emulated_driver_matches_the_fixture(km):
    """tests/golden/kmeans_fixture.npz (the reference's group_paras.py run on the FAISS restatement) against the engine's own
    k-means driver executed under the emulator: same final assignment, same split files — the CPU twin of
    tests/test_gpu_kmeans.py::test_group_paras_flow_on_engine_matches_fixture."""
    import types

    from tests.test_kmeans_oracle import load_kmeans_fixture, run_group_paras_flow
    fx = load_kmeans_fixture()

    class Index:
        def __init__(self, d):
            self.d = d

        def reset(self):
            pass

        def add(self, c):                          # group_paras.py:49-50 re-adds the trained centroids
            c = np.ascontiguousarray(c, np.float32)
            assert km.emu_km_set_centroids(len(c), c.ctypes.data, 0) is None

        def search(self, x, k):
            D, I = emu_assign(km, np.ascontiguousarray(x, np.float32))
            return D[:, None], I[:, None]

    class Clus:
        def __init__(self, d, k):
            self.k, self.niter, self.max_points_per_centroid, self.verbose = k, 25, 256, False

        def train(self, x, index):
            self.centroids, _ = emu_train(km, np.ascontiguousarray(x, np.float32), self.k, self.niter, metric=1, max_pts=self.max_points_per_centroid)
            self.centroids = self.centroids.reshape(-1)

    mod = types.SimpleNamespace(IndexFlatL2=Index, Clustering=Clus, vector_float_to_array=lambda v: np.array(v, dtype=np.float32))
    D, I, samples = run_group_paras_flow(mod, fx)
    np.testing.assert_array_equal(I, fx["I"])
    np.testing.assert_allclose(D, fx["D"], rtol=1e-3, atol=3e-3)
    assert [len(s) for s in samples] == fx["split_sizes"].tolist()
    assert np.concatenate([np.array(s, np.int32) for s in samples]).tolist() == fx["split_lines"].tolist()


# ---- the engine's own stable radix sort (replaces cub::DeviceRadixSort on the k-means path) ------------------------------------
@pytest.mark.parametrize("n,k", [(1, 2), (300, 7), (4096, 256), (4097, 300), (20_000, 10_000), (13_000, 70_000)])
def test_stable_sort_by_centroid_matches_numpy(km, n, k):
    """Pairs (centroid of point i, i) sorted by centroid, stably: every centroid's points come out in ascending point index — the
    order Clustering::train adds them in.  One, two and three 8-bit passes; chunks of 4096 elements per block, ragged last chunk."""
    rng = np.random.default_rng(n + k)
    keys = rng.integers(0, k, n).astype(np.int32)
    keys[: min(n, 50)] = keys[0]                      # a run of equal keys across warp and tile boundaries
    vals = np.arange(n, dtype=np.int32)
    out = np.full(n, -9, np.int32)
    key_bits = 1
    while (1 << key_bits) < k:
        key_bits += 1
    msg = km.emu_km_sort(keys.ctypes.data, vals.ctypes.data, n, key_bits, out.ctypes.data)
    assert msg is None, msg
    np.testing.assert_array_equal(out, np.argsort(keys, kind="stable").astype(np.int32))
