"""-m gpu, OPT-IN (PROQA_B200_STAGED_KMEANS=1): k-means with the points sharded over GPUs (proqa_b200/sharded_clustering.py on
pq_kmeans_set_centroids / _partial_device / _finish_device), one process, against the single-GPU Clustering.

Skipped by default: the staged entry points were written after round 1's GPU budget was spent.  Their logic runs on the CPU
under the SIMT emulator (tests/test_simt_kmeans.py, including the gloo multi-process case); this file is their first
hardware check:  PROQA_B200_STAGED_KMEANS=1 python -m pytest tests/test_gpu_sharded_kmeans.py -m gpu -x -q
(and under torchrun with 2 ranks through tools/gpu_runs/r02_sharded_kmeans.py)."""
import os

import numpy as np
import pytest

from tests.test_kmeans_oracle import blobs

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("PROQA_B200_STAGED_KMEANS") != "1", reason="staged k-means is opt-in until validated on hardware")]


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
