"""-m gpu: k-means with the points sharded over GPUs (proqa_b200/sharded_clustering.py on pq_kmeans_set_centroids /
_partial_device / _finish_device), one process, against the single-GPU Clustering (validated on a B200 in round 2; the two-rank
run under torchrun is tests/workers/sharded_kmeans.py)."""
import numpy as np
import pytest

from tests.test_kmeans_oracle import blobs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("metric,spherical,n,k,mpc", [(1, False, 6000, 20, 1000), (0, True, 6000, 20, 1000), (1, False, 40000, 16, 100)])
def test_staged_steps_on_one_gpu_equal_the_single_driver(metric, spherical, n, k, mpc):
    import proqa_b200 as pq
    from proqa_b200.sharded_clustering import ShardedClustering
    x, _ = blobs(n, k, distinct_init=(mpc * k >= n))
    a, b = pq.IndexFlat(128, metric), pq.IndexFlat(128, metric)
    c1 = pq.Clustering(128, k)
    c1.niter, c1.spherical, c1.max_points_per_centroid = 5, spherical, mpc
    c1.train(x, a)
    c2 = ShardedClustering(128, k)
    c2.niter, c2.spherical, c2.max_points_per_centroid = 5, spherical, mpc
    c2.train(x, b)
    np.testing.assert_array_equal(c2.centroids.view(np.uint32), c1.centroids.view(np.uint32))   # one shard: the very same sums
    np.testing.assert_allclose(c2.obj, c1.obj, rtol=1e-6)
    np.testing.assert_array_equal(a.search(x, 1)[1], b.search(x, 1)[1])
