"""-m gpu: one process, several GPUs (proqa_b200/multi.py on proqa_b200/csrc/pq_multi.cu) — the index the faiss shim hands the
unmodified scripts when PROQA_B200_DEVICES names more than one device (eval_retrieval.py:102-104, group_paras.py:36-51).

With one GPU in the box the "devices" are two shards on device 0 driven by two host threads: the same code path — per-index
locks, concurrent streams, mailboxes written by one shard's kernels while the other shard's kernels read them, gather and
merge — only the NVLink hop is missing.  With two or more GPUs the shards sit on different devices."""
import numpy as np
import pytest

from oracle import oracle
from tests import data

pytestmark = pytest.mark.gpu


def _devices(n=2):
    import torch
    have = torch.cuda.device_count()
    return [g % have for g in range(n)]


def _exact(D, I, xq, xb, k, metric):
    Dr, Ir = oracle.engine_spec(xq, xb, k, metric)
    np.testing.assert_array_equal(I, Ir)
    np.testing.assert_array_equal(D.view(np.uint32), Dr.view(np.uint32))


@pytest.mark.parametrize("metric,n_dev,nq,k", [(0, 2, 40, 100), (1, 3, 24, 10), (0, 4, 9, 1500)])
def test_rows_sharded_over_devices_exact(metric, n_dev, nq, k):
    import proqa_b200 as pq
    xb, xq = data.corpus(1_100_000 + 77), data.queries(nq)          # > 2**20 rows: sharded; a ragged last shard
    ix = pq.MultiGpuIndexFlat(128, metric, _devices(n_dev))
    ix.add(xb)
    assert ix.layout == "rows sharded" and ix.ntotal == len(xb)
    for _ in range(2):                                               # twice: the second search reuses mailboxes and workspaces
        D, I = ix.search(xq, k)
        _exact(D, I, xq, xb, k, metric)
    if k <= 1024:
        assert ix.last_stats[3] > 0, "the tensor-core tier did not run"


def test_threshold_exchange_happens_and_can_be_switched_off(monkeypatch):
    import proqa_b200 as pq
    xb, xq = data.corpus(1_100_000), data.queries(64)
    ix = pq.MultiGpuIndexFlat(128, 0, _devices(2))
    ix.add(xb)
    D, I = ix.search(xq, 80)
    _exact(D, I, xq, xb, 80, 0)
    exchanges = ix.last_stats[9]
    monkeypatch.setenv("PROQA_B200_SHARE", "0")
    iy = pq.MultiGpuIndexFlat(128, 0, _devices(2))
    iy.add(xb)
    D2, I2 = iy.search(xq, 80)
    np.testing.assert_array_equal(I2, I)
    np.testing.assert_array_equal(D2.view(np.uint32), D.view(np.uint32))
    assert iy.last_stats[9] == 0
    assert exchanges >= 0        # (how many arrive in time depends on the schedule; results never do)


def test_rows_in_document_order_sharded():
    """Topic clusters stored contiguously: the shards differ — one holds the rows a query wants, the others learn it."""
    import proqa_b200 as pq
    rng = np.random.default_rng(8)
    cent = rng.standard_normal((40, 128)).astype(np.float32)
    lab = np.sort(rng.integers(0, 40, 1_060_000))
    xb = (cent[lab] + 1.5 * rng.standard_normal((len(lab), 128))).astype(np.float32)
    xq = (cent[rng.integers(0, 40, 48)] + 0.4 * rng.standard_normal((48, 128))).astype(np.float32)
    ix = pq.MultiGpuIndexFlat(128, 0, _devices(4))
    ix.add(xb)
    D, I = ix.search(xq, 100)
    _exact(D, I, xq, xb, 100, 0)


def test_several_adds_keep_insertion_order_ids():
    import proqa_b200 as pq
    xb, xq = data.corpus(1_300_000), data.queries(16)
    ix = pq.MultiGpuIndexFlat(128, 0, _devices(2))
    ix.add(xb[:1_100_000])
    ix.add(xb[1_100_000:1_100_003])      # three rows: the second device's slice of this add is empty or tiny
    ix.add(xb[1_100_003:])
    assert ix.ntotal == len(xb)
    D, I = ix.search(xq, 50)
    Dr, Ir = oracle.engine_spec(xq, xb, 50, 0)
    np.testing.assert_array_equal(D.view(np.uint32), Dr.view(np.uint32))
    np.testing.assert_array_equal(I, Ir)
    ix.reset()
    assert ix.ntotal == 0 and ix.layout == "empty"


def test_small_index_is_replicated_and_queries_are_split():
    """group_paras.py:49-51: index.add(centroids); index.search(data, 1)."""
    import proqa_b200 as pq
    cents, pts = data.corpus(3000), data.queries(200_001)
    for metric in (1, 0):
        ix = pq.MultiGpuIndexFlat(128, metric, _devices(3))
        ix.add(cents)
        assert ix.layout == "rows replicated, queries split"
        D, I = ix.search(pts, 1)
        single = pq.IndexFlat(128, metric)
        single.add(cents)
        Ds, Is = single.search(pts, 1)
        np.testing.assert_array_equal(I, Is)
        np.testing.assert_array_equal(D.view(np.uint32), Ds.view(np.uint32))
        _exact(D[:2000], I[:2000], pts[:2000], cents, 1, metric)


def test_shim_returns_the_multi_index_and_clustering_trains_on_it(monkeypatch):
    import os
    import sys
    import proqa_b200 as pq
    monkeypatch.setenv("PROQA_B200_DEVICES", ",".join(str(d) for d in _devices(2)))
    shim = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "proqa_b200", "faiss_shim")
    monkeypatch.syspath_prepend(shim)
    sys.modules.pop("faiss", None)
    import faiss
    try:
        index = faiss.IndexFlatL2(128)
        assert isinstance(index, pq.MultiGpuIndexFlat)
        from tests.test_kmeans_oracle import blobs
        x, _ = blobs(6000, 20, distinct_init=True)
        clus = faiss.Clustering(128, 20)
        clus.niter, clus.max_points_per_centroid = 5, 1000
        clus.train(x, index)
        ref_ix, ref = pq.IndexFlatL2(128), pq.Clustering(128, 20)
        ref.niter, ref.max_points_per_centroid = 5, 1000
        ref.train(x, ref_ix)
        np.testing.assert_array_equal(clus.centroids.view(np.uint32), ref.centroids.view(np.uint32))
        assert index.ntotal == 20
        np.testing.assert_array_equal(index.search(x, 1)[1], ref_ix.search(x, 1)[1])
        clus.train(x, index)                      # a second train() on the same object starts afresh
        np.testing.assert_array_equal(clus.centroids.view(np.uint32), ref.centroids.view(np.uint32))
    finally:
        sys.modules.pop("faiss", None)
