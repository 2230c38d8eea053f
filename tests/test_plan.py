"""CPU: the launch planner of the tensor-core tier (proqa_b200/csrc/pq_plan.h, through pq_plan_describe — no device needed).
Invariants the kernels rely on, and the plans of the BASELINE.json configurations."""
import ctypes

import numpy as np
import pytest

N_SMS = 148


def _sets():
    """Candidate slabs per row slice = epilogue warp sets of the filter kernel (out[7] of pq_plan_describe)."""
    from proqa_b200 import _lib
    out = (ctypes.c_int64 * 64)()
    assert _lib.lib().pq_plan_describe(100000, 8, 10, N_SMS, out, len(out)) == 0
    return int(out[7])


SETS = _sets()


def plan(ntotal, nq, k, n_sms=N_SMS):
    from proqa_b200 import _lib
    out = (ctypes.c_int64 * (8 + 8 * 64))()
    rc = _lib.lib().pq_plan_describe(ntotal, nq, k, n_sms, out, len(out))
    assert rc == 0, _lib.last_error()
    head = dict(zip(["epochs", "groups", "base", "rem", "m_max", "kp", "nq_pad"], list(out[:7])))
    eps = [dict(zip(["begin", "end", "s1", "s0", "cap", "ctas", "n_sub", "sets"], list(out[8 + 8 * e: 8 + 8 * e + 8]))) for e in range(head["epochs"])]
    return head, eps


CASES = [(21_000_000, 3610, 100), (1_000_000, 2032, 80), (21_000_000, 65536, 80), (12_500_000, 8192, 1000), (10_000, 1 << 20, 1),
         (21_000_000, 16, 80), (21_000_000, 5, 1), (50, 4, 80), (1024, 130, 7), (1025, 600, 1024), (300_000, 262_144, 33), (7, 9, 1)]


@pytest.mark.parametrize("ntotal,nq,k", CASES)
def test_plan_invariants(ntotal, nq, k):
    head, eps = plan(ntotal, nq, k)
    n_mtiles = head["nq_pad"] // 128
    assert head["nq_pad"] >= nq and head["nq_pad"] % 128 == 0 and head["nq_pad"] - nq < 128
    # query tiles are dealt as evenly as possible, never more than 4 per CTA (TMEM: 4 x 64 columns of queries)
    assert head["groups"] * head["base"] + head["rem"] == n_mtiles
    assert 1 <= head["base"] <= head["m_max"] <= 4 and 0 <= head["rem"] < head["groups"]
    assert head["kp"] >= 64 and head["kp"] >= 2.5 * k - 1 and head["kp"] & (head["kp"] - 1) == 0
    # epochs: contiguous, in order, every row exactly once
    assert eps[0]["begin"] == 0 and eps[-1]["end"] == ntotal
    for a, b in zip(eps, eps[1:]):
        assert a["end"] == b["begin"] and a["begin"] < a["end"]
    for i, ep in enumerate(eps):
        tiles = -(-(ep["end"] - ep["begin"]) // 128)
        assert 1 <= ep["s1"] <= tiles and 1 <= ep["s0"] <= tiles, "never more slices than row tiles"
        assert ep["ctas"] == head["rem"] * ep["s1"] + (head["groups"] - head["rem"]) * ep["s0"] >= 1
        assert ep["sets"] in (2, 4) and ep["n_sub"] == ep["sets"] * max(ep["s1"], ep["s0"])
        assert 128 // ep["sets"] <= ep["cap"] <= 4096 and ep["cap"] & (ep["cap"] - 1) == 0
        if ep["begin"] > 0:  # epoch boundaries are multiples of a tile, so a tile never straddles two epochs
            assert ep["begin"] % 128 == 0
        if k > 1 and i == 0:  # bootstrap: every score is a survivor, the slabs must hold a whole row half each
            assert ep["s1"] == ep["s0"] == tiles and ep["sets"] == 4 and ep["cap"] == 32
        if k > 1 and i > 0:   # provision: at least twice the survivors expected on exchangeable rows (1.5 k (end/begin - 1))
            expect = 1.5 * k * (ep["end"] - ep["begin"]) / ep["begin"]
            assert ep["cap"] * ep["sets"] * min(ep["s1"], ep["s0"]) >= min(2 * expect, 4096 * ep["sets"] * min(ep["s1"], ep["s0"]))
        # the candidate slabs of one pass stay far below a B200's 180 GB
        assert head["nq_pad"] * ep["n_sub"] * ep["cap"] * 8 <= 48e9
    if k == 1:
        assert len(eps) == 1


def test_plans_of_the_baseline_configs():
    head, eps = plan(21_000_000, 3610, 100)                  # C2
    assert (head["groups"], head["base"], head["rem"]) == (8, 3, 5) and len(eps) == 6
    assert [e["end"] for e in eps] == [1024, 8192, 65536, 524288, 4194304, 21_000_000]
    assert eps[-1]["ctas"] == 145 and (eps[-1]["s1"], eps[-1]["s0"]) == (20, 15)     # slices in proportion to the tiles owned
    head, eps = plan(21_000_000, 65536, 80)                  # C3: 128 groups on 148 SMs -> 8 slices each, 7 full waves
    assert head["groups"] == 128 and eps[-1]["ctas"] == 1024
    head, eps = plan(10_000, 1 << 20, 1)                     # C4 batch: one pass, one slice per group (79 row tiles only)
    assert len(eps) == 1 and eps[0]["s0"] == 1 and eps[0]["ctas"] == head["groups"] == 2048
    head, eps = plan(21_000_000, 16, 80)                     # small batch: fewer, larger epochs; all SMs on one query tile
    assert len(eps) == 4 and eps[-1]["ctas"] == N_SMS


def test_plan_rejects_bad_arguments():
    from proqa_b200 import _lib
    out = (ctypes.c_int64 * 64)()
    assert _lib.lib().pq_plan_describe(0, 1, 1, N_SMS, out, 64) != 0
    assert _lib.lib().pq_plan_describe(10, 1, 4096, N_SMS, out, 64) != 0      # k above the tensor tier's limit
    assert _lib.lib().pq_plan_describe(21_000_000, 3610, 100, N_SMS, out, 8) != 0   # output too small


# ---- large k (1024 < k <= PQ_MAX_K): sample thresholds + one pass + finalize (pq_plan.h: plan_large_k) ----------------------
LK = ["applies", "step", "k_sample", "sample_rows", "pool", "sort_n", "sample_epochs", "s1", "s0", "cap", "ctas", "n_sub", "batch",
      "slab_bytes", "finalize_smem", "kp_sample"]


def plan_large_k(ntotal, nq, k, n_sms=N_SMS):
    from proqa_b200 import _lib
    out = (ctypes.c_int64 * 16)()
    rc = _lib.lib().pq_plan_describe_large_k(ntotal, nq, k, n_sms, out, len(out))
    assert rc == 0, _lib.last_error()
    return dict(zip(LK, list(out)))


@pytest.mark.parametrize("ntotal,nq,k", [(8_800_000, 6980, 10000), (8_800_000, 500_000, 10000), (21_000_000, 5, 5000), (21_000_000, 128, 15360),
                                          (2_000_000, 64, 1025), (300_000, 300, 2000), (100_000, 16, 10000)])
def test_large_k_plan_invariants(ntotal, nq, k):
    p = plan_large_k(ntotal, nq, k)
    assert p["applies"] == (1 if ntotal >= 64 * k else 0)
    # the sample search is an ordinary k <= 1024 search whose k-th best sits near full-corpus rank 1.35 k
    assert 2 <= p["step"] and 1 <= p["k_sample"] <= 1024
    assert 1.35 * k <= p["k_sample"] * p["step"] <= 1.35 * k + p["step"]
    assert p["sample_rows"] == ntotal // p["step"] and (p["sample_rows"] - 1) * p["step"] + p["step"] // 2 < ntotal
    assert p["kp_sample"] <= 4096
    # finalize kernel: the sort area holds k, the pool holds the rescored set (about 1.5 k), all within one SM's shared memory
    assert p["sort_n"] >= k and p["sort_n"] & (p["sort_n"] - 1) == 0 and p["sort_n"] < 2 * k
    assert p["pool"] >= p["sort_n"] and p["pool"] >= min(2 * k, 24576)
    assert p["finalize_smem"] + 2048 <= 227 * 1024
    # single pass: slabs provisioned for at least 3 x 2.2 k survivors per query, and a batch's slabs stay modest
    assert p["n_sub"] == SETS * max(p["s1"], p["s0"]) and p["cap"] & (p["cap"] - 1) == 0
    assert p["cap"] * SETS * min(p["s1"], p["s0"]) >= 6.6 * k
    assert p["slab_bytes"] <= 16e9
    if p["applies"]:
        assert p["sample_rows"] >= 16 * p["k_sample"] and p["sample_epochs"] >= 2


def test_large_k_plan_of_the_trec_call():
    """retrieval/trec_process.py:76 — index.search(xq, 10000) over the 8.8M MS MARCO passages."""
    p = plan_large_k(8_841_823, 6980, 10000)
    assert (p["step"], p["k_sample"], p["sort_n"], p["pool"]) == (14, 965, 16384, 20000)
    assert p["sample_rows"] == 8_841_823 // 14


def test_large_k_plan_rejects_small_k():
    from proqa_b200 import _lib
    out = (ctypes.c_int64 * 16)()
    assert _lib.lib().pq_plan_describe_large_k(1_000_000, 10, 511, N_SMS, out, 16) != 0       # below 512 the epochs stay
    assert _lib.lib().pq_plan_describe_large_k(1_000_000, 10, 15361, N_SMS, out, 16) != 0


def test_mid_k_plan_of_the_scale_out_config():
    """BASELINE C5: 8192-query batches, k = 1000, 12.5M rows per GPU — thresholds from every 10th row at k_sample = 160."""
    p = plan_large_k(12_500_000, 8192, 1000)
    assert p["applies"] == 1 and (p["step"], p["k_sample"], p["sort_n"]) == (10, 160, 1024) and p["pool"] >= 2 * 1000
    assert p["step"] * p["k_sample"] >= 1.5 * 1000, "the threshold must aim well beyond rank k (rank noise ~ 1/sqrt(k_sample))"
