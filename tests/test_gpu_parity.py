"""-m gpu: the CUDA path (through the C ABI) against the oracle on the same seeded inputs.

Bar (north star): integer/index work bit-exact — every tier must reproduce oracle.engine_spec (the engine's
defined fp32 score and order) bit for bit; and against the fp64 ground truth scores within 1e-4 relative,
ids identical except within that tolerance of a tie.
"""
import numpy as np
import pytest

from oracle import oracle
from tests import data

pytestmark = pytest.mark.gpu


def _index(metric, xb, tier):
    import proqa_b200 as pq
    ix = pq.IndexFlatIP(128) if metric == 0 else pq.IndexFlatL2(128)
    ix.set_tier(tier)
    if len(xb):
        ix.add(xb)
    return ix


def _assert_bit_exact(D, I, Dr, Ir):
    np.testing.assert_array_equal(I, Ir)
    np.testing.assert_array_equal(D.view(np.uint32), Dr.view(np.uint32))


@pytest.mark.parametrize("metric", [0, 1])
@pytest.mark.parametrize("nq,nb,k", [(1, 1000, 10), (3, 5000, 80), (16, 20000, 100), (19, 4097, 1), (21, 12345, 80),
                                      (40, 777, 128), (5, 300, 1000), (2, 20000, 1000)])
def test_fp32_tier_bit_exact(metric, nq, nb, k):
    xb, xq = data.corpus(nb), data.queries(nq)
    ix = _index(metric, xb, "fp32")
    D, I = ix.search(xq, k)
    Dr, Ir = oracle.engine_spec(xq, xb, k, metric)
    _assert_bit_exact(D, I, Dr, Ir)
    assert not oracle.check_against_truth(D, I, xq, xb, k, metric)


@pytest.mark.parametrize("nq,nb,k,kind", [(128, 40000, 80, "normal"), (300, 100000, 100, "normal"), (64, 65536, 1, "normal"),
                                           (200, 50000, 10, "fp16"), (130, 30000, 80, "unit"), (100, 60000, 80, "skewed"),
                                           (512, 20000, 256, "normal"), (1000, 33000, 5, "normal")])
def test_bf16_tier_bit_exact(nq, nb, k, kind):
    xb, xq = data.corpus(nb, kind=kind), data.queries(nq, kind=kind)
    ix = _index(0, xb, "bf16")
    D, I = ix.search(xq, k)
    Dr, Ir = oracle.engine_spec(xq, xb, k, 0)
    _assert_bit_exact(D, I, Dr, Ir)
    st = ix.last_stats
    assert st[0] + st[1] == nq
    assert st[3] > 0, "tensor-core filter did not run"


@pytest.mark.parametrize("nq,nb,k,kind", [(300, 40000, 10, "normal"), (1000, 20000, 1, "normal"), (130, 30000, 80, "skewed"),
                                           (128, 50000, 100, "fp16"), (600, 16500, 1, "unit")])
def test_bf16_tier_l2_bit_exact(nq, nb, k, kind):
    """IndexFlatL2 (group_paras.py:38, the k-means default) on the tensor-core tier: ranking score 2<q,x> - |x|^2."""
    xb, xq = data.corpus(nb, kind=kind), data.queries(nq, kind=kind)
    ix = _index(1, xb, "bf16")
    D, I = ix.search(xq, k)
    Dr, Ir = oracle.engine_spec(xq, xb, k, 1)
    _assert_bit_exact(D, I, Dr, Ir)
    st = ix.last_stats
    assert st[3] > 0, "tensor-core filter did not run"
    assert not oracle.check_against_truth(D[:64], I[:64], xq[:64], xb, k, 1)


def test_tiers_agree_and_match_truth():
    xb, xq = data.corpus(200000), data.queries(256)
    res = {}
    for tier in ("fp32", "bf16", "auto"):
        ix = _index(0, xb, tier)
        res[tier] = ix.search(xq, 80)
        del ix
    _assert_bit_exact(*res["fp32"], *res["bf16"])
    _assert_bit_exact(*res["fp32"], *res["auto"])
    assert not oracle.check_against_truth(*res["auto"], xq, xb, 80, 0)


def test_faiss_restatement_agrees_within_tolerance():
    """What FAISS would print (oracle restatement) vs the engine: same ids except near-ties, scores within 1e-4."""
    xb, xq = data.corpus(50000), data.queries(64)
    ix = _index(0, xb, "auto")
    D, I = ix.search(xq, 80)
    fo = oracle.IndexFlatIP(128)
    fo.add(xb)
    Df, If = fo.search(xq, 80)
    np.testing.assert_allclose(D, Df, rtol=1e-4, atol=1e-4)
    assert (I == If).mean() > 0.999
    assert not oracle.check_against_truth(Df, If, xq, xb, 80, 0)


@pytest.mark.parametrize("tier", ["fp32", "bf16"])
def test_k_larger_than_ntotal_pads(tier):
    xb, xq = data.corpus(50), data.queries(4)
    ix = _index(0, xb, tier)
    D, I = ix.search(xq, 80)
    assert (I[:, 50:] == -1).all() and (D[:, 50:] == -oracle.FLT_MAX).all()
    Dr, Ir = oracle.engine_spec(xq, xb, 80, 0)
    _assert_bit_exact(D, I, Dr, Ir)


def test_empty_index_and_empty_queries():
    import proqa_b200 as pq
    ix = pq.IndexFlatIP(128)
    D, I = ix.search(data.queries(3), 5)
    assert (I == -1).all() and (D == -oracle.FLT_MAX).all()
    ix.add(data.corpus(100))
    D, I = ix.search(np.zeros((0, 128), np.float32), 5)
    assert D.shape == (0, 5) and I.shape == (0, 5)


@pytest.mark.parametrize("tier", ["fp32", "bf16"])
def test_exact_ties_resolve_to_lowest_ids(tier):
    """Duplicated rows straddling the k-th place and the 5/10/20/50 recall cut-offs (eval_retrieval.py:59-64)."""
    base = data.corpus(20000)
    xq = data.queries(32)
    xb = base.copy()
    S = xq[:1] @ base.T
    top = np.argsort(-S[0])[:60]
    for j, t in enumerate(top[:30]):       # every one of the best 30 rows gets 3 exact copies at higher ids
        xb[10000 + 3 * j: 10000 + 3 * j + 3] = base[t]
    ix = _index(0, xb, tier)
    D, I = ix.search(xq, 80)
    Dr, Ir = oracle.engine_spec(xq, xb, 80, 0)
    _assert_bit_exact(D, I, Dr, Ir)
    # within a run of equal scores ids ascend
    for q in range(len(xq)):
        same = D[q, 1:] == D[q, :-1]
        assert (I[q, 1:][same] > I[q, :-1][same]).all()


@pytest.mark.parametrize("tier", ["fp32", "bf16"])
def test_zero_queries_and_zero_rows(tier):
    xb = data.corpus(30000)
    xb[5:9] = 0
    xq = data.queries(16)
    xq[3] = 0
    ix = _index(0, xb, tier)
    D, I = ix.search(xq, 10)
    Dr, Ir = oracle.engine_spec(xq, xb, 10, 0)
    _assert_bit_exact(D, I, Dr, Ir)
    assert I[3].tolist() == list(range(10))  # all-equal scores: FAISS keeps the first k ids


@pytest.mark.parametrize("tier", ["fp32", "auto"])
def test_nan_inf_rows_never_enter(tier):
    xb = data.corpus(20000)
    xb[7, 3] = np.nan
    xb[11, 0] = np.inf
    xb[13, 0] = -np.inf
    xq = np.abs(data.queries(12))
    ix = _index(0, xb, tier)
    D, I = ix.search(xq, 20)
    assert 7 not in I and 13 not in I
    assert np.isfinite(D[:, 1:]).all()
    Dr, Ir = oracle.engine_spec(xq, xb, 20, 0)
    np.testing.assert_array_equal(I[:, 1:], Ir[:, 1:])


def test_repeated_add_reset_and_id_order():
    import proqa_b200 as pq
    xb = data.corpus(9000)
    ix = pq.IndexFlatIP(128)
    ix.set_tier("fp32")
    for a, b in [(0, 1), (1, 1000), (1000, 1001), (1001, 9000)]:
        ix.add(xb[a:b])
    assert ix.ntotal == 9000
    xq = data.queries(7)
    D, I = ix.search(xq, 33)
    _assert_bit_exact(D, I, *oracle.engine_spec(xq, xb, 33, 0))
    ix.reset()
    assert ix.ntotal == 0
    ix.add(xb[:500])
    D, I = ix.search(xq, 33)
    _assert_bit_exact(D, I, *oracle.engine_spec(xq, xb[:500], 33, 0))


def test_kmeans_assignment_shape_l2_and_ip():
    """group_paras.py:49-51 — many points against few centroids, k = 1, both metrics."""
    cents = data.corpus(1000, seed=7)
    pts = data.queries(20000, seed=8)
    for metric in (0, 1):
        Dr, Ir = oracle.engine_spec(pts, cents, 1, metric)
        for tier in ("auto", "bf16"):
            ix = _index(metric, cents, tier)
            D, I = ix.search(pts, 1)
            _assert_bit_exact(D, I, Dr, Ir)


def test_id_base_offsets_ids():
    xb, xq = data.corpus(3000), data.queries(5)
    ix = _index(0, xb, "fp32")
    ix.set_id_base(10_000_000_000)
    D, I = ix.search(xq, 8)
    _, Ir = oracle.engine_spec(xq, xb, 8, 0, id_base=10_000_000_000)
    np.testing.assert_array_equal(I, Ir)


def test_large_k_trec_shape():
    """trec_process.py:76 searches with k = 10000."""
    xb, xq = data.corpus(30000), data.queries(2)
    ix = _index(0, xb, "auto")
    D, I = ix.search(xq, 10000)
    _assert_bit_exact(D, I, *oracle.engine_spec(xq, xb, 10000, 0))


def test_trec_fixture_ids_and_recall_line():
    """tests/golden/trec_fixture.npz: the reference's retrieve_topk() (trec_process.py:69-94, k = 10000) run on the FAISS
    restatement.  The engine must print the same recall line and name the same rows up to fp32 near-ties."""
    from tests.golden_util import assert_same_up_to_near_ties, load_trec_fixture, trec_recall_line
    fx = load_trec_fixture()
    ix = _index(0, fx["xb"], "auto")
    D, I = ix.search(fx["xq"], fx["k"])
    _assert_bit_exact(D, I, *oracle.engine_spec(fx["xq"], fx["xb"], fx["k"], 0))
    assert_same_up_to_near_ties(I, fx["I"], fx["xq"], fx["xb"])
    assert trec_recall_line(I, fx) == fx["recall_line"]


# ---- golden vectors (tests/golden/, generated from the reference's own eval_retrieval.py) ------------------------
@pytest.mark.parametrize("tier", ["fp32", "bf16", "auto"])
def test_eval_fixture_ids_and_recall_lines(tier):
    """eval_retrieval.py:98-123 on the committed fixture: same I as the reference script obtained, and therefore the
    same five 'Top k Recall' lines, byte for byte."""
    from tests.golden_util import load_eval_fixture, recall_lines
    fx = load_eval_fixture()
    ix = _index(0, fx["xb"], tier)
    D, I = ix.search(fx["xq"], fx["topk"])
    np.testing.assert_array_equal(I, fx["I"])
    np.testing.assert_allclose(D, fx["D"], rtol=1e-4, atol=0)
    assert recall_lines(I, fx) == fx["recall_lines"]


@pytest.mark.parametrize("tier", ["fp32", "bf16"])
def test_known_answers_on_engine(tier):
    import json
    import os
    from tests.test_oracle import corpus_from_rule
    cases = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "known_answers.json")))
    for case in cases:
        if tier == "bf16" and case["metric"] == 1:
            continue  # L2 is served by the fp32 scan in every tier
        xb = corpus_from_rule(case)
        ix = _index(case["metric"], xb, tier)
        D, I = ix.search(np.asarray(case["xq"], np.float32), case["k"])
        np.testing.assert_array_equal(D, np.asarray(case["D"], np.float32), err_msg=case["name"])
        want = case.get("I", case.get("I_set"))
        np.testing.assert_array_equal(I, np.asarray(want, np.int64), err_msg=case["name"])  # engine order: id ascending in ties


def test_certificate_failure_falls_back_to_exact_scan():
    """One row with a gigantic norm blows up the error bound E = eps*|q|*max|x|: thresholds stay far below every score, the
    candidate slabs overflow, the certificate fails — and those queries are re-run by the fp32 scan.  Results stay exact."""
    xb, xq = data.corpus(60000), data.queries(40)
    xb[123] *= 3.0e4
    ix = _index(0, xb, "bf16")
    D, I = ix.search(xq, 50)
    Dr, Ir = oracle.engine_spec(xq, xb, 50, 0)
    _assert_bit_exact(D, I, Dr, Ir)
    st = ix.last_stats
    assert st[0] + st[1] == 40
    assert st[1] > 0, "expected the certificate to fail for at least one query (fp32 re-run path not exercised)"


def test_online_sampler_shape_nq1_large_k():
    """qa/online_sampler.py:113 searches one question at a time with k = 5000."""
    xb, xq = data.corpus(100000), data.queries(1)
    ix = _index(0, xb, "auto")
    D, I = ix.search(xq, 5000)
    _assert_bit_exact(D, I, *oracle.engine_spec(xq, xb, 5000, 0))


def test_add_float16_rows_matches_host_widening():
    """index.add(fp16 array) == index.add(fp16.astype(float32)) (eval_retrieval.py:100), through the staged device path too."""
    xh = data.corpus(300000, kind="fp16").astype(np.float16)     # 77 MB: more than one staging chunk
    xq = data.queries(9)
    a, b = _index(0, np.zeros((0, 128), np.float32), "auto"), _index(0, np.zeros((0, 128), np.float32), "auto")
    a.add(xh)
    b.add(xh.astype(np.float32))
    assert a.ntotal == b.ntotal == 300000
    Da, Ia = a.search(xq, 20)
    Db, Ib = b.search(xq, 20)
    _assert_bit_exact(Da, Ia, Db, Ib)


@pytest.mark.parametrize("ncl,nq,spread,expect_second_attempts", [(150, 200, 0.7, False), (8, 600, 2.0, True)])
def test_corpus_in_document_order_stays_on_the_tensor_tier(ncl, nq, spread, expect_second_attempts):
    """Real paragraph indexes are stored in document order: the rows that match a query sit together, far from the prefix
    the first thresholds come from, and enter all at once.  The slabs interleave row tiles and an overflowing epoch is run a
    second time with the threshold its first attempt produced — no query may need the fp32 re-scan, results stay exact.
    (8 clusters of ~37k rows: more survivors in one epoch than the slabs of a query hold.  With tightly packed clusters the
    K'-truncation certificate itself can fail — hundreds of rows within 2E of the k-th — and the fp32 re-scan takes over; the
    wide spread used here keeps the top of each cluster separable, as retrieval scores are.)"""
    rng = np.random.default_rng(11)
    n = 300000
    cent = rng.standard_normal((ncl, 128)).astype(np.float32)
    lab = np.sort(rng.integers(0, ncl, n))                                  # topic-sorted corpus
    xb = (cent[lab] + spread * rng.standard_normal((n, 128))).astype(np.float32)
    qcl = rng.integers(0, ncl, nq)
    xq = (cent[qcl] + 0.3 * rng.standard_normal((nq, 128))).astype(np.float32)
    ix = _index(0, xb, "bf16")
    D, I = ix.search(xq, 100)
    Dr, Ir = oracle.engine_spec(xq, xb, 100, 0)
    _assert_bit_exact(D, I, Dr, Ir)
    st = ix.last_stats
    assert st[1] == 0, f"{st[1]} queries fell back to the fp32 scan (second attempts: {st[8]})"
    if expect_second_attempts:
        assert st[8] > 0, "the second-attempt path was not exercised"


def test_add_npy_streams_float32_and_float16_files(tmp_path):
    xh = data.corpus(5000, kind="fp16")
    np.save(tmp_path / "f32.npy", xh)
    np.save(tmp_path / "f16.npy", xh.astype(np.float16))
    xq = data.queries(6)
    ref = _index(0, xh, "auto").search(xq, 10)
    for name in ("f32.npy", "f16.npy"):
        ix = _index(0, np.zeros((0, 128), np.float32), "auto")
        assert ix.add_npy(str(tmp_path / name), chunk_rows=1500) == 5000 and ix.ntotal == 5000
        _assert_bit_exact(*ix.search(xq, 10), *ref)


def test_many_waves_of_ctas_paced_wave_by_wave():
    """20,000 queries = 40 CTA groups; the last epoch (rows 64k..400k, 2613 row tiles) runs 585 CTAs — four waves on 148 SMs — with
    the TMA producers paced cohort by cohort (pq_mma.cu: pace_blocks_for; BASELINE C3's 65,536 queries run seven such waves).
    Pacing is rate control only: results stay bit-exact."""
    xb, xq = data.corpus(400_000), data.queries(20_000)
    ix = _index(0, xb, "auto")
    D, I = ix.search(xq, 10)
    st = ix.last_stats
    assert st[3] >= 4 and st[1] == 0, st
    sample = np.arange(0, 20_000, 487)                             # 42 queries spread over the groups
    Dr, Ir = oracle.engine_spec(xq[sample], xb, 10, 0)
    _assert_bit_exact(D[sample], I[sample], Dr, Ir)


def test_large_query_and_result_arrays_take_the_staged_copies():
    """pq_index_search moves host arrays of 8 MB or more through the pinned double buffer (pq_index.cu: staged_upload /
    staged_download; group_paras.py:51 hands index.search 10.75 GB of points, C5 returns 98 MB).  20,000 queries = 10 MB up;
    k = 100: D 8 MB and I 16 MB down (odd sizes, so no copy ends on a round boundary); the result must be the same bits as small
    searches of the same queries, which take the plain copies."""
    xb, xq = data.corpus(120_000), data.queries(20_011)
    ix = _index(0, xb, "auto")
    D, I = ix.search(xq, 100)
    assert D.shape == (20_011, 100) and I.shape == (20_011, 100)
    for lo in (0, 9_973, 20_011 - 7):                              # slices small enough for the plain-copy path
        Ds, Is = ix.search(xq[lo:lo + 7], 100)
        _assert_bit_exact(D[lo:lo + 7], I[lo:lo + 7], Ds, Is)
    sample = np.array([0, 4_999, 20_010])
    _assert_bit_exact(D[sample], I[sample], *oracle.engine_spec(xq[sample], xb, 100, 0))
