"""-m gpu: the tensor-core path for 1024 < k <= PQ_MAX_K (pq_mma_largek.inl; the default for such k since round 2) against the
oracle — the retrieval/trec_process.py:76 (k = 10000) and qa/online_sampler.py:113 (k = 5000) call sites."""
import numpy as np
import pytest

from oracle import oracle
from tests import data

pytestmark = pytest.mark.gpu


def _search(metric, xb, xq, k):
    import proqa_b200 as pq
    ix = pq.IndexFlatIP(128) if metric == 0 else pq.IndexFlatL2(128)
    ix.add(xb)
    D, I = ix.search(xq, k)
    return D, I, ix.last_stats


@pytest.mark.parametrize("metric,nq,nb,k,kind", [(0, 64, 200_000, 2000, "normal"), (0, 130, 300_000, 1025, "fp16"), (1, 40, 260_000, 4000, "normal"),
                                                  (0, 5, 700_000, 10000, "normal"), (0, 300, 150_000, 2048, "skewed")])
def test_large_k_tensor_tier_bit_exact(metric, nq, nb, k, kind):
    xb, xq = data.corpus(nb, kind=kind), data.queries(nq, kind=kind)
    D, I, st = _search(metric, xb, xq, k)
    Dr, Ir = oracle.engine_spec(xq, xb, k, metric)
    np.testing.assert_array_equal(I, Ir)
    np.testing.assert_array_equal(D.view(np.uint32), Dr.view(np.uint32))
    assert st[0] + st[1] == nq and st[3] > 0, "tensor-core filter did not run"
    assert st[1] <= nq // 10, f"{st[1]} of {nq} queries fell back to the fp32 scan"


def test_large_k_on_rows_in_document_order():
    """Topic clusters stored contiguously: the row-strided sample still lands the threshold near rank 1.35 k."""
    rng = np.random.default_rng(5)
    centres = rng.standard_normal((400, 128)).astype(np.float32)
    xb = np.concatenate([c + 0.7 * rng.standard_normal((int(rng.integers(50, 1500)), 128)).astype(np.float32) for c in centres])
    xq = (centres[rng.integers(0, 400, 48)] + 0.5 * rng.standard_normal((48, 128))).astype(np.float32)
    k = 1500
    assert len(xb) >= 64 * k
    D, I, st = _search(0, xb, xq, k)
    Dr, Ir = oracle.engine_spec(xq, xb, k, 0)
    np.testing.assert_array_equal(I, Ir)
    np.testing.assert_array_equal(D.view(np.uint32), Dr.view(np.uint32))
    assert st[3] > 0


def test_small_corpus_stays_on_the_scan():
    xb, xq = data.corpus(50_000), data.queries(16)
    D, I, st = _search(0, xb, xq, 1500)          # 50k < 64 k rows: no sample to speak of
    Dr, Ir = oracle.engine_spec(xq, xb, 1500, 0)
    np.testing.assert_array_equal(I, Ir)
    assert st[3] == 0
