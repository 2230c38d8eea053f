"""CPU: the FAISS Clustering restatement (oracle) — what group_paras.py:40-47 relies on."""
import numpy as np

from oracle import oracle
from tests import data


def blobs(n, k, seed=5, spread=0.05, distinct_init=False):
    """k well-separated blobs.  With distinct_init the k points FAISS picks as initial centroids (the first k entries of
    rand_perm(n, seed 1234 + 1)) come from k different blobs, so every blob owns exactly one centroid from the first
    iteration on and every assignment is unambiguous for any fp32 accumulation order."""
    rng = np.random.default_rng(seed)
    centers = rng.standard_normal((k, 128)).astype(np.float32) * 4.0
    lab = rng.integers(0, k, size=n)
    if distinct_init:
        lab[oracle.rand_perm(n, 1235)[:k]] = np.arange(k)
    x = centers[lab] + spread * rng.standard_normal((n, 128)).astype(np.float32)
    return np.ascontiguousarray(x, np.float32), lab


def test_rand_perm_is_a_permutation_and_seeded():
    p = oracle.rand_perm(1000, 1234)
    assert sorted(p.tolist()) == list(range(1000))
    assert (p == oracle.rand_perm(1000, 1234)).all()
    assert (p != oracle.rand_perm(1000, 1235)).any()
    # std::mt19937(1234) first outputs are 822569775, 2137449171, ... : perm[0] = 0 + 822569775 % 1000
    assert p[0] == 822569775 % 1000


def test_kmeans_recovers_blobs_l2_and_objective_decreases():
    x, lab = blobs(6000, 20)
    for metric in (oracle.METRIC_L2,):
        ix = oracle.FaissFlatOracle(128, metric)
        clus = oracle.FaissClusteringOracle(128, 20)
        clus.niter = 8
        clus.train(x, ix)
        assert ix.ntotal == 20
        assert (np.diff(clus.obj) <= 1e-3 * np.abs(clus.obj[:-1])).all(), clus.obj
        D, I = ix.search(x, 1)
        # every blob maps to one centroid (up to clusters the random init merged/split)
        purity = np.mean([np.bincount(I[lab == b, 0]).max() / (lab == b).sum() for b in range(20)])
        assert purity > 0.8


def test_kmeans_void_clusters_are_split():
    # 3 distinct points repeated: with k = 8 most clusters start empty after the first assignment
    base = data.corpus(3, seed=9)
    x = np.repeat(base, 200, axis=0)
    ix = oracle.FaissFlatOracle(128, oracle.METRIC_L2)
    clus = oracle.FaissClusteringOracle(128, 8)
    clus.niter = 3
    clus.train(x, ix)
    assert clus.nsplit[0] > 0
    assert np.isfinite(clus.centroids).all()


def test_kmeans_subsamples_when_too_many_points():
    x, _ = blobs(3000, 4)
    ix = oracle.FaissFlatOracle(128, oracle.METRIC_L2)
    clus = oracle.FaissClusteringOracle(128, 4)
    clus.niter, clus.max_points_per_centroid = 4, 100     # 400 of 3000 points are used
    clus.train(x, ix)
    ix2 = oracle.FaissFlatOracle(128, oracle.METRIC_L2)
    clus2 = oracle.FaissClusteringOracle(128, 4)
    clus2.niter, clus2.max_points_per_centroid = 4, 100
    perm = oracle.rand_perm(3000, 1234)[:400]
    clus2.train(x[perm], ix2)                              # the same 400 points given directly: no subsampling
    np.testing.assert_array_equal(clus.centroids, clus2.centroids)


def test_kmeans_nx_equals_k_copies_points():
    x = data.corpus(16)
    ix = oracle.FaissFlatOracle(128, oracle.METRIC_L2)
    clus = oracle.FaissClusteringOracle(128, 16)
    clus.train(x, ix)
    np.testing.assert_array_equal(clus.centroids.reshape(16, 128), x)


def load_kmeans_fixture():
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kmeans_fixture.npz"))
    return dict(x=np.float32(z["x"]), k=int(z["k"]), niter=int(z["niter"]), mpc=int(z["max_points_per_centroid"]), I=z["I"].astype(np.int64),
                D=z["D"], split_sizes=z["split_sizes"], split_lines=z["split_lines"])


def run_group_paras_flow(faiss_like, fx):
    """group_paras.py:20-53 + :12-18 restated against any module with the faiss surface the script uses."""
    x = fx["x"]
    index = faiss_like.IndexFlatL2(x.shape[1])
    clus = faiss_like.Clustering(x.shape[1], fx["k"])
    clus.verbose = False
    clus.niter = fx["niter"]
    clus.max_points_per_centroid = fx["mpc"]
    clus.train(x, index)
    centroids = faiss_like.vector_float_to_array(clus.centroids).reshape(fx["k"], x.shape[1])
    index.reset()
    index.add(centroids)
    D, I = index.search(x, 1)
    samples = [[] for _ in range(fx["k"])]
    for i in range(len(x)):
        samples[I[i][0]].append(i)
    return D, I, samples


def test_group_paras_fixture_reproduced_by_the_oracle():
    import types
    fx = load_kmeans_fixture()
    mod = types.SimpleNamespace(IndexFlatL2=oracle.IndexFlatL2, Clustering=oracle.FaissClusteringOracle,
                                vector_float_to_array=lambda v: np.array(v, dtype=np.float32))
    D, I, samples = run_group_paras_flow(mod, fx)
    np.testing.assert_array_equal(I, fx["I"])
    assert [len(s) for s in samples] == fx["split_sizes"].tolist()
    assert np.concatenate([np.array(s, np.int32) for s in samples]).tolist() == fx["split_lines"].tolist()
