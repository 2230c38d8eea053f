"""CPU: the exact fp32 tier — pq_ffma_scan_kernel<QT> (TMA-staged streaming scan with the fused shared-memory top-k), its
launcher, search_fp32_scan and pq_merge_lists_kernel — executed under the SIMT emulator and checked bit for bit against the
oracle, the CPU twin of tests/test_gpu_parity.py::test_fp32_tier_bit_exact.  TMA and mbarriers are stand-ins (the box copy in
the SWIZZLE_128B shared-memory layout, a phase counter); everything else is the engine's own source text.  This is the tier
that answers nq <= 4, k > 1024 and every query whose tensor-tier certificate fails.  Test infrastructure only."""
import numpy as np
import pytest

from oracle import oracle
from tests import data
from tests.simt import harness

pytestmark = pytest.mark.timeout(600)   # an emulated kernel that never finishes must not hang the suite


@pytest.fixture(scope="module")
def scan(tmp_path_factory):
    return harness.build_fp32_scan_emu(tmp_path_factory.mktemp("simt_scan"))


def _exact(D, I, Dr, Ir):
    np.testing.assert_array_equal(I, Ir)
    np.testing.assert_array_equal(D.view(np.uint32), Dr.view(np.uint32))


@pytest.mark.parametrize("metric", [0, 1])
@pytest.mark.parametrize("nq,nb,k,n_sms", [(1, 1000, 10, 6), (3, 5000, 80, 6), (8, 4097, 100, 3), (11, 3000, 1, 148), (2, 777, 128, 2),
                                           (5, 300, 1000, 6), (2, 6000, 1000, 4)])
def test_fp32_scan_bit_exact(scan, metric, nq, nb, k, n_sms):
    xb, xq = data.corpus(nb), data.queries(nq)
    D, I, st = harness.run_fp32_scan_emu(scan, xb, xq, k, metric, n_sms=n_sms)
    _exact(D, I, *oracle.engine_spec(xq, xb, k, metric))
    assert st[2] == 1 and st[4] == 1              # query batches of up to 8, all in ONE scan launch and one merge launch


def test_large_k_runs_one_query_per_batch(scan):
    """trec_process.py:76 / online_sampler.py:113 shapes: k in the thousands — the shared-memory buffers leave room for one
    query per batch (CTA) and a single TMA stage."""
    xb, xq = data.corpus(12_000), data.queries(2)
    D, I, st = harness.run_fp32_scan_emu(scan, xb, xq, 10000, 0, n_sms=3)
    _exact(D, I, *oracle.engine_spec(xq, xb, 10000, 0))
    assert st[2] == 1 and st[4] == 1              # (two batches of one query, one launch)


def test_ties_padding_id_base_and_non_finite_rows(scan):
    xb = data.corpus(500)
    xb[100:110] = xb[7]                       # exact ties: lowest ids first
    xb[3, 0] = np.nan
    xb[5, 0] = -np.inf
    xq = np.concatenate([xb[7:8], np.abs(data.queries(2))])
    D, I, _ = harness.run_fp32_scan_emu(scan, xb, xq, 600, 0, id_base=1_000_000)      # k > ntotal: padded
    Dr, Ir = oracle.engine_spec(xq, xb, 600, 0, id_base=1_000_000)
    _exact(D, I, Dr, Ir)
    assert I[0, :4].tolist() == [1_000_007, 1_000_100, 1_000_101, 1_000_102]
    assert (I[:, -2:] == -1).all() and 1_000_003 not in I and 1_000_005 not in I


@pytest.mark.parametrize("schedule", [1, 2])
def test_scan_does_not_depend_on_the_thread_schedule(scan, schedule):
    xb, xq = data.corpus(9000), data.queries(6)
    try:
        D, I, _ = harness.run_fp32_scan_emu(scan, xb, xq, 50, 1, n_sms=5, schedule=schedule)
    finally:
        scan.emu_set_schedule(0)
    _exact(D, I, *oracle.engine_spec(xq, xb, 50, 1))


def test_many_query_batches_share_one_launch(scan):
    """The k-means re-run shape: many queries (here 37: five batches, the last one short) against a small corpus, k = 1 — one scan
    launch whose CTAs are dealt to the batches, one merge launch, results bit-exact."""
    xb, xq = data.corpus(2_000), data.queries(37)
    for metric in (0, 1):
        D, I, st = harness.run_fp32_scan_emu(scan, xb, xq, 1, metric, n_sms=4)
        _exact(D, I, *oracle.engine_spec(xq, xb, 1, metric))
        assert st[2] == 1 and st[4] == 1
