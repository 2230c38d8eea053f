"""Small end-to-end exercise of every kernel, meant to run under compute-sanitizer (developer tool)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import proqa_b200 as pq  # noqa: E402

rng = np.random.default_rng(0)
xb = rng.standard_normal((20000, 128), dtype=np.float32)
for metric in (0, 1):
    ix = pq.IndexFlat(128, metric)
    ix.add(xb)
    for tier, nq, k in (("fp32", 3, 10), ("bf16", 300, 10), ("bf16", 130, 1), ("bf16", 600, 100)):
        ix.set_tier(tier)
        D, I = ix.search(rng.standard_normal((nq, 128), dtype=np.float32), k)
        assert (I >= 0).all()
    del ix
x = rng.standard_normal((3000, 128), dtype=np.float32)
ix = pq.IndexFlatL2(128)
clus = pq.Clustering(128, 16)
clus.niter = 2
clus.train(x, ix)
print("sanitize_small: done")

# ---- paths written after round 1's last GPU run (opt-in): PROQA_B200_SANITIZE_NEW=1 ---------------------------------------
if os.environ.get("PROQA_B200_SANITIZE_NEW") == "1":
    import ctypes

    import torch

    from proqa_b200 import _lib
    from proqa_b200.sharded_clustering import ShardedClustering
    os.environ["PROQA_B200_LARGEK"] = "1"                   # read when an index is created
    big = rng.standard_normal((100000, 128), dtype=np.float32)
    for metric in (0, 1):
        ix = pq.IndexFlat(128, metric)
        ix.add(big)
        D, I = ix.search(rng.standard_normal((9, 128), dtype=np.float32), 1500)    # sample thresholds + one pass + finalize
        assert (I >= 0).all() and ix.last_stats[3] > 0
        del ix
    ix = pq.IndexFlatL2(128)
    sc = ShardedClustering(128, 16)                         # staged k-means steps, one rank
    sc.niter = 2
    sc.train(x, ix)
    G, nq, k = 2, 3, 10000                                  # shard merge by ranking (20000 keys per query)
    D_all = torch.sort(torch.randn(G, nq, k, device="cuda"), dim=2, descending=True).values.contiguous()
    I_all = (torch.arange(k, device="cuda").view(1, 1, k) + torch.arange(G, device="cuda").view(G, 1, 1) * k).expand(G, nq, k).contiguous()
    D_out, I_out = torch.empty(nq, k, device="cuda"), torch.empty(nq, k, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    rc = _lib.lib().pq_merge_shard_results(0, 0, G, nq, k, ctypes.c_void_p(D_all.data_ptr()), ctypes.c_void_p(I_all.data_ptr()),
                                           ctypes.c_void_p(D_out.data_ptr()), ctypes.c_void_p(I_out.data_ptr()))
    _lib.check(rc, "merge")
    assert bool((D_out[:, :-1] >= D_out[:, 1:]).all())
    print("sanitize_small: new paths done")
