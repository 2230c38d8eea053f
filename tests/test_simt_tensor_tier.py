"""CPU: the host driver of the default tensor-core tier, pq::search_mma_filter (k <= 1024: epochs, candidate slabs, carry
lists, second attempts, rescoring + certificate, the k = 1 path), executed on the CPU and checked bit for bit against the
oracle — the same checks tests/test_gpu_parity.py makes on the B200, minus the tcgen05 kernel itself.

Real: the driver's source, the planner (pq_plan.h), and pq_mma_init_state_kernel / pq_epoch_select_kernel / pq_rescore_kernel
/ pq_k1_finalize_kernel under the SIMT emulator (tests/simt/simt_emu.h).  Stand-in: pq_mma_filter_kernel, replaced by a
functional model with the same CTA mapping, slab layout and admission rules (tests/simt/mma_host_emu.cpp.in).  This guards
the host logic and the selection kernels against regressions on machines without a GPU; it is test infrastructure only.
"""
import numpy as np
import pytest

from oracle import oracle
from tests import data
from tests.simt import harness

pytestmark = pytest.mark.timeout(600)   # an emulated kernel that never finishes must not hang the suite


@pytest.fixture(scope="module", params=["warp_per_query", "cta_per_query"])
def host_emu(request, tmp_path_factory):
    """Both selection / rescoring kernel families: one warp per query (what large batches use: K' <= 256 and >= 1024 queries)
    and one CTA per query (small batches, large K')."""
    import os
    old = os.environ.get("PROQA_B200_SELECT_WARP_MIN")
    os.environ["PROQA_B200_SELECT_WARP_MIN"] = "1" if request.param == "warp_per_query" else "1000000000"
    lib = harness.build_host_emu(tmp_path_factory.mktemp("simt_host_" + request.param))
    lib.family = request.param
    yield lib
    if old is None:
        os.environ.pop("PROQA_B200_SELECT_WARP_MIN", None)
    else:
        os.environ["PROQA_B200_SELECT_WARP_MIN"] = old


def _warp_family_only(lib):
    """K' > 256, k = 1 and the schedule fuzzing do not depend on the family (or always take the CTA kernels): run them once."""
    if lib.family != "warp_per_query":
        pytest.skip("covered by the warp-per-query run (this case does not depend on the selection-kernel family)")


def _exact(D, I, xq, xb, k, metric, rows=None):
    rows = list(range(len(xq))) if rows is None else rows
    Dr, Ir = oracle.engine_spec(xq[rows], xb, k, metric)
    np.testing.assert_array_equal(I[rows], Ir)
    np.testing.assert_array_equal(D[rows].view(np.uint32), Dr.view(np.uint32))


@pytest.mark.parametrize("metric,nb,nq,k,kind,n_sms", [(0, 40_000, 20, 80, "normal", 8), (1, 30_000, 8, 10, "normal", 8), (0, 20_000, 130, 256, "fp16", 6),
                                                        (0, 9_000, 5, 1024, "normal", 4), (0, 25_000, 12, 100, "skewed", 148)])
def test_epoch_search_matches_the_oracle(host_emu, metric, nb, nq, k, kind, n_sms):
    xb, xq = data.corpus(nb, kind=kind), data.queries(nq, kind=kind)
    D, I, rerun, st = harness.run_host_emu(host_emu, xb, xq, k, metric, n_sms)
    assert rerun == []
    _exact(D, I, xq, xb, k, metric, rows=list(range(min(nq, 12))))
    assert st[3] >= 2 and st[4] == st[3] + 1     # one select per epoch, then the rescoring kernel


@pytest.mark.parametrize("metric,nb,nq", [(1, 2_000, 300), (0, 1_500, 200), (1, 130, 140)])
def test_k1_assignment_path_matches_the_oracle(host_emu, metric, nb, nq):
    """group_paras.py:45,51 — few centroids (the rows), many points (the queries), k = 1."""
    _warp_family_only(host_emu)
    xb, xq = data.corpus(nb), data.queries(nq)
    D, I, rerun, st = harness.run_host_emu(host_emu, xb, xq, 1, metric)
    assert rerun == []
    _exact(D, I, xq, xb, 1, metric)
    assert st[3] == 1 and st[4] == 1             # one filter pass, one finalize


def test_k_larger_than_ntotal_pads(host_emu):
    _warp_family_only(host_emu)
    xb, xq = data.corpus(50), data.queries(4)
    D, I, rerun, _ = harness.run_host_emu(host_emu, xb, xq, 80, 0)
    assert rerun == []
    assert (I[:, 50:] == -1).all() and (D[:, 50:] == -oracle.FLT_MAX).all()
    _exact(D, I, xq, xb, 80, 0)


def test_rows_in_document_order_take_the_second_attempt(host_emu):
    """Topic-sorted rows: a whole cluster beats the threshold inside one epoch, the slabs overflow, the epoch is run again
    with the threshold the first attempt produced (pq_epoch_select_kernel: allow_redo / is_redo)."""
    rng = np.random.default_rng(11)
    n, ncl, nq, k = 120_000, 2, 6, 100
    cent = rng.standard_normal((ncl, 128)).astype(np.float32)
    lab = np.sort(rng.integers(0, ncl, n))
    xb = (cent[lab] + 2.0 * rng.standard_normal((n, 128))).astype(np.float32)
    xq = (cent[rng.integers(0, ncl, nq)] + 0.3 * rng.standard_normal((nq, 128))).astype(np.float32)
    D, I, rerun, st = harness.run_host_emu(host_emu, xb, xq, k, 0, n_sms=4)
    assert st[8] > 0, "the second-attempt path was not exercised"
    ok = [q for q in range(nq) if q not in rerun]
    assert len(ok) >= nq - 1
    _exact(D, I, xq, xb, k, 0, rows=ok)


@pytest.mark.parametrize("schedule", [1, 2])
def test_results_do_not_depend_on_the_thread_schedule(host_emu, schedule):
    """The emulator resumes a block's threads in a different pseudo-random order after every barrier / warp collective, so
    that another thread runs ahead each time; shared state read after a barrier that a thread running ahead may already have
    changed shows up as wrong results or as a divergent barrier (tests/test_simt_select.py has the worked example)."""
    if host_emu.family != "warp_per_query" and schedule == 2:
        pytest.skip("one fuzzed schedule per family")
    try:
        for metric, nb, nq, k in ((0, 30_000, 10, 80), (1, 1_500, 150, 1), (0, 80_000, 5, 1100)):
            xb, xq = data.corpus(nb), data.queries(nq)
            D, I, rerun, _ = harness.run_host_emu(host_emu, xb, xq, k, metric, schedule=schedule)
            assert rerun == []
            _exact(D, I, xq, xb, k, metric, rows=list(range(min(nq, 6))))
    finally:
        host_emu.emu_set_schedule(0)


# ---- corpus row-sharded over several GPUs: threshold exchange through peer mailboxes (pq_mma.cu: ShareParams) ----------------
def _merge_lists(Ds, Is, k, metric):
    """numpy statement of pq_merge_shard_results: best first, ties to the lower id, -1 padding last."""
    D = np.concatenate(Ds, axis=1)
    I = np.concatenate(Is, axis=1)
    Do, Io = np.empty((len(D), k), np.float32), np.empty((len(D), k), np.int64)
    for q in range(len(D)):
        valid = I[q] >= 0
        order = np.lexsort((I[q], -D[q] if metric == 0 else D[q], ~valid))[:k]
        Do[q], Io[q] = D[q, order], I[q, order]
        pad = ~valid[order]
        Do[q, pad] = oracle.FLT_MAX if metric == 1 else -oracle.FLT_MAX
        Io[q, pad] = -1
    return Do, Io


@pytest.mark.parametrize("metric,nb,nq,k,R,ordered", [(0, 60_000, 9, 80, 2, False), (0, 90_000, 6, 100, 3, False), (1, 60_000, 7, 20, 4, False),
                                                       (0, 80_000, 6, 100, 2, True), (0, 40_000, 5, 256, 2, False)])
def test_row_shards_exchanging_thresholds_merge_to_the_exact_result(host_emu, metric, nb, nq, k, R, ordered):
    """Every shard filters at what the shards know TOGETHER (max of the k-th best, min of the ceil(k/R)-th best local scores):
    its list may hold fewer than k rows, but the merge of the lists is the exact global top-k.  The emulator runs the shards
    one after the other, so in the first round shard r sees the final values of shards < r only (and nothing of the others:
    the bounded wait runs out); in the second round (same sequence number) it sees every other shard's final values next to
    its own early ones — the k-th-best rule bites when the shards differ (rows in document order), the ceil(k/R) rule needs
    the shards in step, which only real GPUs give (tests/test_gpu_sharded.py)."""
    if host_emu.family != "warp_per_query" and (R != 2 or ordered):
        pytest.skip("the CTA-per-query family publishes through the same code: two of the cases")
    if ordered:
        rng = np.random.default_rng(21)
        cent = rng.standard_normal((3, 128)).astype(np.float32)
        lab = np.sort(rng.integers(0, 3, nb))
        xb = (cent[lab] + 2.0 * rng.standard_normal((nb, 128))).astype(np.float32)
        xq = (cent[rng.integers(0, 3, nq)] + 0.3 * rng.standard_normal((nq, 128))).astype(np.float32)
    else:
        xb, xq = data.corpus(nb), data.queries(nq)
    Dr, Ir = oracle.engine_spec(xq, xb, k, metric)
    bound = (float(harness.engine_norms(xb).max()), float(harness.resid2(xb).max()))     # maxima over the WHOLE corpus
    cap_q = 64
    boxes = [np.zeros(R * cap_q * 2 + R, np.uint64) for _ in range(R)]
    per = (nb + R - 1) // R
    tightened = 0
    for rnd in range(2):
        Ds, Is = [], []
        for r in range(R):
            lo, hi = r * per, min(nb, (r + 1) * per)
            D, I, rerun, st = harness.run_host_emu(host_emu, xb[lo:hi], xq, k, metric, n_sms=6,
                                                   share=(R, r, cap_q, 3, 77, lo, boxes), bound=bound)
            assert rerun == []
            Ds.append(D)
            Is.append(I)
            tightened += int((I == -1).any())
        D, I = _merge_lists(Ds, Is, k, metric)
        np.testing.assert_array_equal(I, Ir)
        np.testing.assert_array_equal(D.view(np.uint32), Dr.view(np.uint32))
    if ordered:
        assert tightened > 0, "no shard ever used a peer's threshold (lists were all complete local top-k lists)"
    # values of another search (a different sequence number) are ignored: a lone shard then returns its full local top-k
    D, I, rerun, _ = harness.run_host_emu(host_emu, xb[:per], xq, k, metric, n_sms=6, share=(R, 0, cap_q, 3, 78, 0, boxes), bound=bound)
    Dl, Il = oracle.engine_spec(xq, xb[:per], k, metric)
    np.testing.assert_array_equal(I, Il)
