"""-m gpu: parity at the sizes BASELINE.json is quoted on ("GPU parity tests never touch a BASELINE size").

C1 (configs[0], eval_retrieval.py shape): 2032 queries x 1M x 128, k = 80 — bit-exact against the oracle's engine_spec on a query
sample, and the whole batch inside the north-star tolerance of the fp64 ground truth (oracle.check_against_truth).
C2 (configs[1]): 3610 queries x 21M x 128, k = 100 — corpus generated on the device (the 10.75 GB never exist on the host), every
returned (score, id) of a 128-query sample checked against fp64 brute force: the score is the fp64 score of THAT id within
1e-4 relative, and the ids' fp64 scores are the true top-k scores rank by rank (ties apart)."""
import os
import sys

import numpy as np
import pytest

from oracle import oracle
from tests import data

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c1_eval_retrieval_shape():
    import proqa_b200 as pq
    nq, N, k = 2032, 1_000_000, 80
    xb, xq = data.corpus(N), data.queries(nq)
    ix = pq.IndexFlatIP(128)
    ix.add(xb)
    D, I = ix.search(xq, k)
    assert ix.last_stats[3] > 0 and ix.last_stats[1] == 0          # tensor-core tier, no query re-run by the scan
    sample = np.arange(0, nq, 43)                                  # 48 queries, bit for bit
    Dr, Ir = oracle.engine_spec(xq[sample], xb, k, 0)
    np.testing.assert_array_equal(I[sample], Ir)
    np.testing.assert_array_equal(D[sample].view(np.uint32), Dr.view(np.uint32))
    chk = np.arange(0, nq, 8)                                      # 254 queries against the fp64 truth, north-star tolerance
    assert not oracle.check_against_truth(D[chk], I[chk], xq[chk], xb, k, 0)
    # and the same through the exact fp32 scan: one defined score, whatever the tier
    ix.set_tier("fp32")
    D2, I2 = ix.search(xq[:8], k)
    np.testing.assert_array_equal(I2, I[:8])
    np.testing.assert_array_equal(D2.view(np.uint32), D[:8].view(np.uint32))


def test_c2_nq_scale_shape_device_generated():
    import torch
    sys.path.insert(0, ROOT)
    import bench
    import proqa_b200 as pq
    nq, N, k = 3610, 21_000_000, 100
    dev = torch.device("cuda", 0)
    ix = pq.IndexFlatIP(128, 0)
    bench.build_shard(ix, 0, N, dev, n_global=N)
    assert ix.ntotal == N
    xq = bench.host_queries(nq)
    D, I = ix.search(xq, k)                                        # the call eval_retrieval.py:104 makes
    assert ix.last_stats[3] > 0 and ix.last_stats[1] == 0
    assert (I >= 0).all() and (I < N).all()
    assert (np.diff(D, axis=1) <= 0).all(), "rows must be best-first"
    nchk = 128
    q = torch.from_numpy(xq[:nchk]).to(dev)
    Dt, It, own = bench.truth_topk_fp64(q, 0, N, N, k, dev, ids=torch.from_numpy(I[:nchk]).to(dev))
    ok, info = bench.parity_gate(torch.from_numpy(D[:nchk]).to(dev), torch.from_numpy(I[:nchk]).to(dev), Dt, It, own)
    assert ok, info
    assert info["ids_equal_frac"] > 0.999 and info["max_near_tie_gap_rel"] < 1e-4
    # Small batches on the same 21M-row index (S0; online_sampler.py:113 asks one question at a time, k = 5000): from 2^23 rows on
    # AUTO sends them to the tensor tier too (half the bytes of the fp32 scan), and the two tiers return the same bits.
    for nq_s, k_s in ((1, 80), (4, 80), (1, 5000), (3, 1000)):
        ix.set_tier("auto")
        Da, Ia = ix.search(xq[:nq_s], k_s)
        st = ix.last_stats
        assert st[3] > 0 and st[1] == 0, f"nq={nq_s} k={k_s}: expected the tensor tier without re-runs, stats {st}"
        ix.set_tier("fp32")
        Df, If = ix.search(xq[:nq_s], k_s)
        assert ix.last_stats[3] == 0
        np.testing.assert_array_equal(Ia, If)
        np.testing.assert_array_equal(Da.view(np.uint32), Df.view(np.uint32))
    np.testing.assert_array_equal(Ia[:, :100], I[:3])               # (and a prefix of the big batch's answer)
