"""CPU: pq_mma_filter_kernel — the tcgen05 / TMEM / TMA kernel of the tensor-core tier — executed ITSELF under the SIMT emulator,
inside the real host driver, and checked bit for bit against the oracle.

tests/test_simt_tensor_tier.py replaces this kernel by a functional model; here its own source runs: the three warp roles (TMA
producer, MMA issuer, eight epilogue warps), the mbarrier protocol with its phase parities, the rotation of the two TMEM
accumulators, operand staging, the L2 bias epilogue, the k = 1 running maximum, the second-attempt path.  What the hardware does
is a model written down in tests/simt/filter_tcgen05.inc (mbarrier = phase + arrivals + bytes; TMA box copies with the 128-byte
swizzle; TMEM as 128 x 512 words with lane-quarter access rules; tcgen05.mma decoded from its descriptors).  The model can be
wrong about the silicon — the B200 tests settle that — but given the model, a protocol slip (a barrier initialised for the
wrong number of arrivals, a consumer two phases ahead, a warp reading another warp's TMEM quarter) shows up here, on the CPU, as
an error or a reported deadlock instead of a hung GPU.  Test infrastructure only."""
import numpy as np
import pytest

from oracle import oracle
from tests import data
from tests.simt import harness

pytestmark = pytest.mark.timeout(600)   # an emulated kernel that never finishes must not hang the suite


@pytest.fixture(scope="module")
def real(tmp_path_factory):
    return harness.build_host_emu(tmp_path_factory.mktemp("simt_filter"), real_filter=True)


def _exact(D, I, xq, xb, k, metric, rows):
    Dr, Ir = oracle.engine_spec(xq[rows], xb, k, metric)
    np.testing.assert_array_equal(I[rows], Ir)
    np.testing.assert_array_equal(D[rows].view(np.uint32), Dr.view(np.uint32))


# nq = 130 / 500 / 700 give CTA groups owning 2 / 4 / 3 query tiles (M_TILES template 2, 4, 4 with m = 3)
@pytest.mark.parametrize("metric,nb,nq,k,kind,n_sms", [(0, 5_000, 7, 10, "normal", 4), (1, 5_000, 7, 10, "normal", 4), (0, 4_133, 130, 80, "fp16", 3),
                                                        (1, 3_000, 130, 1, "normal", 4), (0, 2_500, 200, 1, "normal", 2), (0, 6_000, 500, 5, "normal", 8),
                                                        (1, 3_000, 700, 3, "normal", 5)])
def test_filter_kernel_inside_the_driver_matches_the_oracle(real, metric, nb, nq, k, kind, n_sms):
    xb, xq = data.corpus(nb, kind=kind), data.queries(nq, kind=kind)
    D, I, rerun, st = harness.run_host_emu(real, xb, xq, k, metric, n_sms)
    assert rerun == [] and st[3] >= 1
    _exact(D, I, xq, xb, k, metric, rows=list(range(min(nq, 8))) + [nq - 1])


@pytest.mark.parametrize("schedule", [1, 2])
def test_filter_kernel_under_fuzzed_schedules(real, schedule):
    try:
        for metric, nb, nq, k in ((0, 5_000, 7, 10), (1, 3_000, 130, 1), (1, 4_000, 300, 20)):
            xb, xq = data.corpus(nb), data.queries(nq)
            D, I, rerun, _ = harness.run_host_emu(real, xb, xq, k, metric, 4, schedule=schedule)
            assert rerun == []
            _exact(D, I, xq, xb, k, metric, rows=list(range(min(nq, 6))))
    finally:
        real.emu_set_schedule(0)


def test_second_attempt_path_with_the_real_kernel(real):
    rng = np.random.default_rng(11)
    n, nq, k = 40_000, 4, 100
    cent = rng.standard_normal((2, 128)).astype(np.float32)
    lab = np.sort(rng.integers(0, 2, n))
    xb = (cent[lab] + 2.0 * rng.standard_normal((n, 128))).astype(np.float32)
    xq = (cent[rng.integers(0, 2, nq)] + 0.3 * rng.standard_normal((nq, 128))).astype(np.float32)
    D, I, rerun, st = harness.run_host_emu(real, xb, xq, k, 0, n_sms=2)
    ok = [q for q in range(nq) if q not in rerun]
    assert len(ok) >= nq - 1
    _exact(D, I, xq, xb, k, 0, rows=ok)


@pytest.mark.parametrize("old,new,expect", [
    ("mbar_init(&ctrl->tmem_empty[b], kEpiWarps);", "mbar_init(&ctrl->tmem_empty[b], kEpiWarps + 1);", b"deadlock"),
    ("if (lane == 0) mbar_arrive(&ctrl->tmem_empty[b]);", "", b"deadlock"),
    ("const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(set * kSubN);",
     "const uint32_t taddr0 = tmem_base + ((uint32_t)(set * 32) << 16) + (uint32_t)(set * kSubN);", b"lane quarter"),
])
def test_protocol_slips_are_caught_on_the_cpu(tmp_path, old, new, expect):
    """The same kernel with one line changed: a barrier initialised for one arrival too many, a missing arrival, a warp
    addressing another warp's TMEM lanes."""
    lib = harness.build_host_emu(tmp_path, real_filter=True, mutate=(old, new))
    xb, xq = data.corpus(3_000), data.queries(7)
    try:
        D, I, rerun, _ = harness.run_host_emu(lib, xb, xq, 10, 0, 4)
    except AssertionError as e:
        assert expect is None or expect in str(e).encode(), str(e)
        return
    # the run came back: then the results must be wrong (or every query sent to the scan), never silently right
    Dr, Ir = oracle.engine_spec(xq, xb, 10, 0)
    ok = [q for q in range(7) if q not in rerun]
    assert len(ok) < 7 or not np.array_equal(I, Ir), "a protocol slip went unnoticed"


@pytest.mark.parametrize("mode", [2])   # (mode 1 — single-wave grids only — is the same code with the first branch unpaced)
def test_producer_pacing_counts_every_cta_out_of_every_block(tmp_path, monkeypatch, mode):
    """Pacing of the TMA producers (pq_mma.cu: pace_blocks_for / pace_leave / pace_wait), forced on for a small search: blocks of
    two row tiles.  Results are unchanged (pacing is rate control only), and after each paced launch every CTA of a cohort has
    counted itself out of every block exactly once — a producer that missed one would leave the others waiting on the hardware.
    mode 1: single-wave grids only; mode 2: larger grids pace wave by wave (cohorts of n_sms CTAs)."""
    monkeypatch.setenv("PROQA_B200_PACE", str(mode))
    monkeypatch.setenv("PROQA_B200_PACE_MIN_TILES", "1")
    monkeypatch.setenv("PROQA_B200_PACE_SHIFT", "1")
    lib = harness.build_host_emu(tmp_path, real_filter=True)
    n_sms, nb, nq, k = 5, 3_000, 700, 3            # 6 query tiles -> two CTA groups; 24 row tiles
    xb, xq = data.corpus(nb), data.queries(nq)
    D, I, rerun, st = harness.run_host_emu(lib, xb, xq, k, 1, n_sms)
    assert rerun == []
    _exact(D, I, xq, xb, k, 1, rows=list(range(8)) + [nq - 1])
    import ctypes
    buf = (ctypes.c_uint32 * 64)()
    n = lib.emu_last_pace(buf, 64)
    counts = list(buf)[:min(n, 64)]
    # The bootstrap epoch (8 row tiles: 4 blocks) runs one CTA per row tile and group, 16 CTAs on 5 SMs: unpaced in mode 1, four
    # cohorts of 5, 5, 5 and 1 CTAs in mode 2.  The second epoch has 16 row tiles (8 blocks) and a single-wave grid of 4 CTAs.
    if mode == 1:
        assert counts == [4] * 8, counts
    else:
        assert counts == [5] * 12 + [1] * 4 + [4] * 8, counts
