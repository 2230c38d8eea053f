"""-m gpu: the k-means driver (pq_kmeans_train behind proqa_b200.Clustering) against the FAISS Clustering restatement."""
import numpy as np
import pytest

from oracle import oracle
from tests.test_kmeans_oracle import blobs, load_kmeans_fixture, run_group_paras_flow

pytestmark = pytest.mark.gpu


def _train_both(x, k, niter, metric, spherical=False, mpc=256):
    import proqa_b200 as pq
    ix = pq.IndexFlat(128, metric)
    clus = pq.Clustering(128, k)
    clus.niter, clus.spherical, clus.max_points_per_centroid = niter, spherical, mpc
    clus.train(x, ix)
    io = oracle.FaissFlatOracle(128, metric)
    co = oracle.FaissClusteringOracle(128, k)
    co.niter, co.spherical, co.max_points_per_centroid = niter, spherical, mpc
    co.train(x, io)
    return clus, ix, co, io


@pytest.mark.parametrize("n,k,niter", [(6000, 20, 6), (40000, 64, 4)])
def test_kmeans_l2_blobs_match_faiss_restatement_bit_for_bit(n, k, niter):
    """Well separated blobs: assignments are unambiguous, and the centroid update adds points in index order in fp32 exactly
    as km_update_centroids does, so the whole trajectory (centroids after every iteration) is bit-identical."""
    x, _ = blobs(n, k, distinct_init=True)
    clus, ix, co, io = _train_both(x, k, niter, 1, mpc=1000)   # no subsampling: the init points are rand_perm(n, 1235)[:k]
    np.testing.assert_array_equal(clus.centroids.view(np.uint32), co.centroids.view(np.uint32))
    np.testing.assert_allclose(clus.obj, co.obj, rtol=5e-4)   # FAISS sums |x|^2+|y|^2-2xy in fp32: ~1e-4 absolute noise per distance
    assert ix.ntotal == k
    D, I = ix.search(x, 1)
    Do, Io = io.search(x, 1)
    np.testing.assert_array_equal(I, Io)


def test_kmeans_l2_shared_blobs_stay_close():
    """Random init may drop two centroids into one blob: points near their boundary can flip between implementations
    (fp32 rounding of the distances), so the trajectories agree to tolerance rather than bit for bit."""
    x, _ = blobs(6000, 20)
    clus, ix, co, io = _train_both(x, 20, 6, 1)
    np.testing.assert_allclose(clus.centroids, co.centroids, rtol=0, atol=2e-2)
    np.testing.assert_allclose(clus.obj, co.obj, rtol=1e-3)
    D, I = ix.search(x, 1)
    Do, Io = io.search(x, 1)
    assert (I == Io).mean() > 0.95   # a blob shared by two near-coincident centroids is cut almost arbitrarily


def test_kmeans_subsampling_and_void_split_match():
    base = np.random.default_rng(3).standard_normal((5, 128)).astype(np.float32) * 3
    x = np.repeat(base, 400, axis=0) + 0.01 * np.random.default_rng(4).standard_normal((2000, 128)).astype(np.float32)
    clus, ix, co, io = _train_both(x, 12, 4, 1, mpc=100)    # 1200 of 2000 points; 12 centroids over 5 blobs: splits happen
    np.testing.assert_allclose(clus.obj[0], co.obj[0], rtol=1e-3)    # same subsample, same initial centroids
    np.testing.assert_allclose(clus.obj, co.obj, rtol=0.1)           # later iterations: near-coincident centroids, loose
    assert np.isfinite(clus.centroids).all()


def test_kmeans_spherical_ip_runs_and_normalises():
    x, _ = blobs(5000, 16)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    clus, ix, co, io = _train_both(x, 16, 5, 0, spherical=True)
    c = clus.centroids.reshape(16, 128)
    np.testing.assert_allclose(np.linalg.norm(c, axis=1), 1.0, rtol=1e-5)
    np.testing.assert_allclose(clus.obj, co.obj, rtol=1e-2)


def test_group_paras_flow_on_engine_matches_fixture():
    """retrieval/group_paras.py:20-53,12-18 on the committed fixture (made by running the reference script on the FAISS
    restatement): same final assignment, same split files."""
    import proqa_b200 as pq
    import types
    fx = load_kmeans_fixture()
    mod = types.SimpleNamespace(IndexFlatL2=pq.IndexFlatL2, Clustering=pq.Clustering, vector_float_to_array=pq.vector_float_to_array)
    D, I, samples = run_group_paras_flow(mod, fx)
    np.testing.assert_array_equal(I, fx["I"])
    np.testing.assert_allclose(D, fx["D"], rtol=1e-3, atol=3e-3)   # |x|^2+|y|^2-2xy in fp32: terms ~2e3, results ~0.3
    assert [len(s) for s in samples] == fx["split_sizes"].tolist()
    assert np.concatenate([np.array(s, np.int32) for s in samples]).tolist() == fx["split_lines"].tolist()


def test_kmeans_errors():
    import proqa_b200 as pq
    ix = pq.IndexFlatL2(128)
    clus = pq.Clustering(128, 50)
    with pytest.raises(RuntimeError, match="should be at least as large as number of clusters"):   # FAISS's exception type and text
        clus.train(np.zeros((10, 128), np.float32), ix)          # fewer points than clusters
    bad = np.zeros((100, 128), np.float32)
    bad[3, 3] = np.nan
    with pytest.raises(RuntimeError, match="NaN"):
        clus.train(bad, ix)
