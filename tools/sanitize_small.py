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
