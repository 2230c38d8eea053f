"""group_paras.py shape: faiss.Clustering(128, 10000).train(x[10M,128]) — seconds per iteration (developer benchmark)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import proqa_b200 as pq  # noqa: E402

n, k, niter = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000, 10_000, 5
g = torch.Generator(device="cuda")
g.manual_seed(1)
x = torch.randn((n, 128), generator=g, device="cuda", dtype=torch.float32).cpu().numpy()
for metric, name in ((pq.METRIC_L2, "L2"), (pq.METRIC_INNER_PRODUCT, "IP/spherical")):
    ix = pq.IndexFlat(128, metric)
    clus = pq.Clustering(128, k)
    clus.niter, clus.max_points_per_centroid, clus.verbose = niter, 1000, True
    clus.spherical = metric == pq.METRIC_INNER_PRODUCT
    t0 = time.perf_counter()
    clus.train(x, ix)
    dt = time.perf_counter() - t0
    print(f"\n[{name}] n={n} k={k} niter={niter}: total {dt:.2f}s, obj={clus.obj.tolist()}", flush=True)
    t0 = time.perf_counter()
    D, I = ix.search(x[:2_000_000], 1)
    print(f"[{name}] final assignment of 2M points through index.search (host in/out): {time.perf_counter()-t0:.3f}s", flush=True)
