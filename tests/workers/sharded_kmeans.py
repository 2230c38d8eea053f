"""Round-2 hardware check of ShardedClustering across ranks:
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/workers/sharded_kmeans.py
Every rank trains on the same array; rank 0 compares with the single-GPU Clustering."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import proqa_b200 as pq  # noqa: E402
from proqa_b200.sharded_clustering import ShardedClustering  # noqa: E402
from tests.test_kmeans_oracle import blobs  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, k, niter = 400_000, 1000, 5
x, _ = blobs(n, k, distinct_init=False)
ix = pq.IndexFlat(128, pq.METRIC_L2, local)
clus = ShardedClustering(128, k)
clus.niter, clus.max_points_per_centroid = niter, 256
t0 = time.perf_counter()
clus.train(x, ix)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
if rank == 0:
    ref_ix = pq.IndexFlat(128, pq.METRIC_L2, local)
    ref = pq.Clustering(128, k)
    ref.niter, ref.max_points_per_centroid = niter, 256
    ref.train(x, ref_ix)
    err = np.abs(clus.centroids - ref.centroids).max()
    same = (ix.search(x[:50000], 1)[1] == ref_ix.search(x[:50000], 1)[1]).mean()
    print(f"sharded k-means: {dt:.2f}s for {niter} iterations on {dist.get_world_size()} ranks; max |centroid diff| {err:.3g}; "
          f"same assignment {same:.4f}; objective {clus.obj.tolist()} vs {ref.obj.tolist()}")
dist.destroy_process_group()
