"""CPU: the selection / merge kernels of pq_select.cu — pq_merge_lists_kernel (fp32 scan: the per-CTA top-k lists of a query ->
its result) and pq_merge_di_kernel (multi-GPU: the R shard results -> one, pq_merge_shard_results) — with their launchers,
executed under the SIMT emulator (tests/simt) and compared with a plain numpy statement of what they must produce.
Test infrastructure only; the kernels' source is taken verbatim from the engine."""
import ctypes

import numpy as np
import pytest

from tests.simt import harness
from tests.simt.harness import f32_ordered

FLT_MAX = np.float32(3.4028234663852886e38)

pytestmark = pytest.mark.timeout(600)   # an emulated kernel that never finishes must not hang the suite


@pytest.fixture(scope="module")
def sel(tmp_path_factory):
    return harness.build_select_emu(tmp_path_factory.mktemp("simt_select"))


def make_keys(scores, rows):
    return (f32_ordered(scores).astype(np.uint64) << np.uint64(32)) | ((~rows.astype(np.uint32)) & np.uint32(0xFFFFFFFF)).astype(np.uint64)


def key_score(keys):
    o = (keys >> np.uint64(32)).astype(np.uint32)
    u = np.where(o & 0x80000000, o & 0x7FFFFFFF, ~o).astype(np.uint32)
    return u.view(np.float32)


def key_row(keys):
    return (~(keys & np.uint64(0xFFFFFFFF)).astype(np.uint32)).astype(np.int64)


@pytest.mark.parametrize("nq,n_lists,list_len,k,metric,with_counts,with_gthr", [(3, 16, 100, 100, 0, False, False), (2, 148, 80, 80, 1, True, True),
                                                                             (2, 40, 1000, 1000, 0, True, False), (1, 7, 30, 50, 0, False, True),
                                                                             (1, 30, 5000, 5000, 0, False, False)])
def test_merge_lists_kernel(sel, nq, n_lists, list_len, k, metric, with_counts, with_gthr):
    rng = np.random.default_rng(nq * 1000 + n_lists)
    rows = rng.permutation(n_lists * list_len * nq * 3)[:nq * n_lists * list_len].reshape(nq, n_lists, list_len)
    scores = rng.standard_normal((nq, n_lists, list_len)).astype(np.float32)
    scores[:, :, ::7] = np.float32(0.25)                       # exact ties: order must fall back to the row id
    keys = make_keys(scores, rows)
    keys[rng.random(keys.shape) < 0.1] = 0                     # empty slots
    counts = rng.integers(0, list_len + 1, (nq, n_lists)).astype(np.uint32) if with_counts else None
    gthr = f32_ordered(np.full(nq, -0.3, np.float32)) if with_gthr else None
    q_norms = (rng.random(nq) * 50 + 100).astype(np.float32)
    D = np.full((nq, k), np.nan, np.float32)
    I = np.full((nq, k), -7, np.int64)
    out_keys = np.zeros((nq, k), np.uint64)
    msg = sel.emu_merge_lists(keys.ctypes.data, n_lists * list_len, list_len, n_lists, list_len, counts.ctypes.data if with_counts else None, n_lists,
                              gthr.ctypes.data if with_gthr else None, nq, k, metric, q_norms.ctypes.data, 1000, D.ctypes.data, I.ctypes.data,
                              out_keys.ctypes.data)
    assert msg is None, msg.decode()
    for q in range(nq):
        live = keys[q].copy()
        if with_counts:
            live[np.arange(list_len)[None, :] >= counts[q][:, None]] = 0
        live = live[live != 0]
        if with_gthr:
            live = live[live >= (np.uint64(gthr[q]) << np.uint64(32))]
        want = np.sort(live)[::-1][:k]
        n = len(want)
        np.testing.assert_array_equal(out_keys[q, :n], want)
        assert (out_keys[q, n:] == 0).all()
        np.testing.assert_array_equal(I[q, :n], key_row(want) + 1000)
        assert (I[q, n:] == -1).all()
        s = key_score(want)
        want_d = np.maximum(np.float32(0), q_norms[q] - s) if metric == 1 else s
        np.testing.assert_array_equal(D[q, :n].view(np.uint32), want_d.astype(np.float32).view(np.uint32))
        assert (D[q, n:] == (FLT_MAX if metric == 1 else -FLT_MAX)).all()


def reference_merge(D_all, I_all, k, metric):
    """Best-first; ties -> the lower global id; -1 padding last (what pq_merge_shard_results documents)."""
    G, nq, _ = D_all.shape
    D = D_all.transpose(1, 0, 2).reshape(nq, -1)
    I = I_all.transpose(1, 0, 2).reshape(nq, -1)
    Do = np.empty((nq, k), np.float32)
    Io = np.empty((nq, k), np.int64)
    for q in range(nq):
        valid = I[q] >= 0
        key = -D[q] if metric == 0 else D[q]
        order = np.lexsort((I[q], key, ~valid))[:k]
        Do[q], Io[q] = D[q, order], I[q, order]
        pad = ~valid[order]
        Do[q, pad] = FLT_MAX if metric == 1 else -FLT_MAX
        Io[q, pad] = -1
    return Do, Io


@pytest.mark.parametrize("G,nq,k,metric", [(2, 5, 20, 0), (8, 3, 100, 1), (4, 2, 1000, 0), (3, 4, 7, 1), (1, 2, 16, 0),
                                           (2, 2, 10000, 0), (8, 1, 4000, 1), (3, 2, 15360, 0)])   # the last three: beyond the in-CTA sort
def test_merge_shard_results_kernel(sel, G, nq, k, metric):
    rng = np.random.default_rng(G * 100 + k)
    per = max(5000, 2 * k)
    D_all = np.empty((G, nq, k), np.float32)
    I_all = np.empty((G, nq, k), np.int64)
    for g in range(G):       # shard g owns ids [g*per, (g+1)*per); its list is best-first with ties in ascending id order
        for q in range(nq):
            n_valid = k if (g + q) % 3 else k // 2          # some shards hold fewer than k rows: -1 padding
            ids = np.sort(rng.choice(per, n_valid, replace=False)) + g * per
            sc = np.round(rng.standard_normal(n_valid), 1).astype(np.float32)      # coarse scores: many ties within and across shards
            if metric == 1:
                sc = np.abs(sc)
            order = np.lexsort((ids, -sc if metric == 0 else sc))
            D_all[g, q, :n_valid], I_all[g, q, :n_valid] = sc[order], ids[order]
            D_all[g, q, n_valid:] = FLT_MAX if metric == 1 else -FLT_MAX
            I_all[g, q, n_valid:] = -1
    D, I = harness.merge_di(sel, D_all, I_all, k, metric)
    Dr, Ir = reference_merge(D_all, I_all, k, metric)
    np.testing.assert_array_equal(I, Ir)
    np.testing.assert_array_equal(D.view(np.uint32), Dr.view(np.uint32))


@pytest.mark.parametrize("G,k", [(2, 64), (2, 10000), (3, 6000)])
def test_merge_l2_equal_distances_with_unsorted_ids(sel, G, k):
    """L2 lists are ordered by the engine's key 2<q,x> - |x|^2 before D = max(0, |q|^2 - key) is rounded, so equal D values can
    carry ids that are not ascending.  Both merge kernels break such ties by (list, position): every output slot is written once."""
    rng = np.random.default_rng(k + G)
    nq = 2
    D_all = np.sort(np.round(np.abs(rng.standard_normal((G, nq, k))), 1).astype(np.float32), axis=2)
    I_all = np.empty((G, nq, k), np.int64)
    for g in range(G):
        for q in range(nq):
            I_all[g, q] = rng.permutation(k) + g * k          # ids in arbitrary order inside runs of equal D
    D, I = harness.merge_di(sel, D_all, I_all, k, 1)
    Dc = D_all.transpose(1, 0, 2).reshape(nq, -1)
    Ic = I_all.transpose(1, 0, 2).reshape(nq, -1)
    for q in range(nq):
        order = np.argsort(Dc[q], kind="stable")[:k]           # stable: (list, position) decides ties
        np.testing.assert_array_equal(I[q], Ic[q, order])
        np.testing.assert_array_equal(D[q], Dc[q, order])


def test_merge_refuses_more_lists_than_the_ranking_kernel_takes(sel):
    G, k = 65, 300                                     # 65 x 300 keys do not fit the in-CTA sort, and 65 lists exceed the ranking kernel
    D_all = np.zeros((G, 1, k), np.float32)
    I_all = np.zeros((G, 1, k), np.int64)
    D = np.empty((1, k), np.float32)
    I = np.empty((1, k), np.int64)
    msg = sel.emu_merge_di(D_all.ctypes.data, I_all.ctypes.data, G, 1, k, 0, D.ctypes.data, I.ctypes.data)
    assert msg is not None and b"refused" in msg


# ---- the barrier after the fill count is read (pq_merge_lists_kernel) -------------------------------------------------------
def _fill_boundary_case():
    """One query, lists of 1024 entries: the first batch leaves 1022 entries in the work array (two empty slots), so the
    'sort now?' test  fill + 1024 > 2048  is false — unless a thread that ran ahead has already appended its candidates of
    the second batch."""
    rng = np.random.default_rng(1)
    n_lists, list_len, k = 3, 1024, 64
    scores = rng.standard_normal((1, n_lists, list_len)).astype(np.float32)
    rows = np.arange(n_lists * list_len).reshape(1, n_lists, list_len)
    keys = make_keys(scores, rows)
    keys[0, 0, 5] = 0
    keys[0, 0, 700] = 0
    return keys, n_lists, list_len, k


def _run_merge(lib, keys, n_lists, list_len, k):
    D = np.empty((1, k), np.float32)
    I = np.empty((1, k), np.int64)
    out = np.zeros((1, k), np.uint64)
    q_norms = np.zeros(1, np.float32)
    msg = lib.emu_merge_lists(keys.ctypes.data, n_lists * list_len, list_len, n_lists, list_len, None, n_lists, None, 1, k, 0, q_norms.ctypes.data, 0,
                              D.ctypes.data, I.ctypes.data, out.ctypes.data)
    return msg, out


def test_fill_count_barrier_holds_under_adversarial_schedules(sel, tmp_path):
    keys, n_lists, list_len, k = _fill_boundary_case()
    want = np.sort(keys[keys != 0])[::-1][:k]
    try:
        for seed in (0, 1, 2, 3, 4, 5):
            sel.emu_set_schedule(seed)
            msg, out = _run_merge(sel, keys, n_lists, list_len, k)
            assert msg is None, (seed, msg)
            np.testing.assert_array_equal(out[0], want)
    finally:
        sel.emu_set_schedule(0)
    # the same kernel without that barrier: some schedule lets a thread run ahead into the next batch before the others have
    # read the fill count, and the block-uniform branch stops being uniform
    broken = harness.build_select_emu(tmp_path, drop="__syncthreads();  // every thread has read the fill before")
    bad = 0
    for seed in (0, 1, 2, 3, 4, 5):
        broken.emu_set_schedule(seed)
        msg, out = _run_merge(broken, keys, n_lists, list_len, k)
        bad += (msg is not None) or not np.array_equal(out[0], want)
    assert bad > 0, "the emulator's schedules no longer expose the hazard this barrier closes"


# ---- add()-time row preparation (pq_prep_rows_kernel): bf16 copy, engine norms, residuals, non-finite flags ------------------
def test_prep_rows_kernel_matches_the_numpy_model(sel):
    from tests import data
    x = data.corpus(700, kind="skewed")
    x[13, 5] = np.inf
    x[14, 9] = np.nan
    x[15, 1] = 3.39e38                        # finite in fp32, overflows bf16
    n = len(x)
    bf = np.zeros((n, 128), np.uint16)
    norms = np.zeros(n, np.float32)
    row_bad = np.full(n, 7, np.uint8)
    resid = np.zeros(n, np.float32)
    scal = np.zeros(3, np.uint32)             # max norm bits, non-finite flag, max residual bits
    msg = sel.emu_prep_rows(x.ctypes.data, n, bf.ctypes.data, norms.ctypes.data, scal[0:].ctypes.data, scal[1:].ctypes.data, row_bad.ctypes.data,
                            resid.ctypes.data, scal[2:].ctypes.data)
    assert msg is None, msg
    good = np.ones(n, bool)
    good[[13, 14, 15]] = False
    np.testing.assert_array_equal(bf[good], harness.bf16_bits(x[good]))                          # round to nearest even
    np.testing.assert_array_equal(norms[good].view(np.uint32), harness.engine_norms(x[good]).view(np.uint32))
    assert row_bad[good].sum() == 0 and row_bad[[13, 14, 15]].tolist() == [1, 1, 1] and scal[1] == 1
    model = (((x[good] - harness.bf16_round(x[good])).astype(np.float64) ** 2).sum(1))
    assert (resid[good] >= model * 0.99999).all() and (resid[good] <= model * 1.001 + 1e-30).all()   # an upper bound, tight
    assert scal[2:].view(np.float32)[0] == resid[good].max()
    assert scal[0:1].view(np.float32)[0] >= norms[good].max()
