"""-m gpu: the exchange + merge over peer memory (proqa_b200/csrc/pq_xchg.cu) through the C ABI, against numpy.

The ranks of a job are played by W exchange objects in ONE process on device 0, each on its own CUDA stream, their buffers
connected by plain device pointers: the kernels, flags, slice geometry and double-buffered result areas are the real ones, only
the NVLink hop is missing (tests/test_gpu_sharded.py runs the two-process / two-GPU version under torchrun)."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
FLT_MAX = np.float32(np.finfo(np.float32).max)


def _bounds(n, parts, i):
    per = (n + parts - 1) // parts
    lo = min(n, i * per)
    return lo, min(n, lo + per)


def _local_lists(R, Q, nq, k, metric, seed):
    """lists[rank] = (D, I) of rank = rq * R + rr: the best-first list of row shard rr for the queries of group rq."""
    rng = np.random.default_rng(seed)
    per_rows = max(5000, 2 * k)
    out = []
    for rq in range(Q):
        qlo, qhi = _bounds(nq, Q, rq)
        for rr in range(R):
            n = qhi - qlo
            D = np.empty((n, k), np.float32)
            I = np.empty((n, k), np.int64)
            for q in range(n):
                n_valid = k if (rr + q) % 3 else k // 2
                ids = np.sort(rng.choice(per_rows, n_valid, replace=False)) + rr * per_rows
                sc = np.round(rng.standard_normal(n_valid), 1).astype(np.float32) + np.float32(0.0)
                if metric == 1:
                    sc = np.abs(sc)
                order = np.lexsort((ids, -sc if metric == 0 else sc))
                D[q, :n_valid], I[q, :n_valid] = sc[order], ids[order]
                D[q, n_valid:] = FLT_MAX if metric == 1 else -FLT_MAX
                I[q, n_valid:] = -1
            out.append((D, I))
    return out


def _reference(lists, R, Q, nq, k, metric):
    Do, Io = np.empty((nq, k), np.float32), np.empty((nq, k), np.int64)
    for rq in range(Q):
        qlo, qhi = _bounds(nq, Q, rq)
        for q in range(qhi - qlo):
            D = np.concatenate([lists[rq * R + rr][0][q] for rr in range(R)])
            I = np.concatenate([lists[rq * R + rr][1][q] for rr in range(R)])
            valid = I >= 0
            order = np.lexsort((I, -D if metric == 0 else D, ~valid))[:k]
            Do[qlo + q], Io[qlo + q] = D[order], I[order]
            pad = ~valid[order]
            Do[qlo + q, pad] = FLT_MAX if metric == 1 else -FLT_MAX
            Io[qlo + q, pad] = -1
    return Do, Io


@pytest.mark.parametrize("R,Q,nq,k,metric", [(2, 1, 37, 100, 0), (4, 1, 50, 10, 1), (1, 4, 41, 80, 0), (2, 2, 29, 33, 0), (3, 1, 7, 1, 1),
                                             (2, 1, 5, 10000, 0), (8, 1, 64, 80, 0)])
def test_exchange_and_merge_match_numpy(R, Q, nq, k, metric):
    import torch
    from proqa_b200 import _lib
    L = _lib.lib()
    W = R * Q
    need = L.pq_xchg_bytes_needed(nq, k, R, Q)
    xs, bases = [], (ctypes.c_void_p * W)()
    for r in range(W):
        h, base, nbytes = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int64()
        _lib.check(L.pq_xchg_create(0, W, r, need, ctypes.byref(h), ctypes.byref(base), ctypes.byref(nbytes)), "xchg_create")
        xs.append(h)
        bases[r] = base.value
    for r in range(W):
        _lib.check(L.pq_xchg_connect(xs[r], bases), "xchg_connect")
    streams = [torch.cuda.Stream() for _ in range(W)]
    try:
        for seq in (1, 2, 3):                                   # consecutive searches: flags and result areas are reused
            lists = _local_lists(R, Q, nq, k, metric, 1000 * seq + R * 10 + Q)
            Dr, Ir = _reference(lists, R, Q, nq, k, metric)
            dev_in = [(torch.from_numpy(D).cuda(), torch.from_numpy(I).cuda()) for D, I in lists]
            dev_out = [(torch.full((nq, k), float("nan"), device="cuda"), torch.full((nq, k), -7, dtype=torch.int64, device="cuda")) for _ in range(W)]
            torch.cuda.synchronize()
            for r in range(W):
                Dl, Il = dev_in[r]
                rc = L.pq_xchg_run(xs[r], metric, R, nq, k, ctypes.c_void_p(Dl.data_ptr()), ctypes.c_void_p(Il.data_ptr()),
                                   ctypes.c_void_p(dev_out[r][0].data_ptr()), ctypes.c_void_p(dev_out[r][1].data_ptr()), seq,
                                   ctypes.c_void_p(streams[r].cuda_stream))
                _lib.check(rc, "xchg_run")
            torch.cuda.synchronize()
            for r in range(W):
                _lib.check(L.pq_xchg_check(xs[r]), "xchg_check")
                np.testing.assert_array_equal(dev_out[r][1].cpu().numpy(), Ir, err_msg=f"ids on rank {r}, search {seq}")
                np.testing.assert_array_equal(dev_out[r][0].cpu().numpy().view(np.uint32), Dr.view(np.uint32), err_msg=f"scores on rank {r}")
    finally:
        for h in xs:
            L.pq_xchg_free(h)


def test_buffers_too_small_are_refused():
    import torch
    from proqa_b200 import _lib
    L = _lib.lib()
    h, base, nbytes = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int64()
    _lib.check(L.pq_xchg_create(0, 1, 0, L.pq_xchg_bytes_needed(10, 10, 1, 1), ctypes.byref(h), ctypes.byref(base), ctypes.byref(nbytes)), "create")
    arr = (ctypes.c_void_p * 1)(base.value)
    _lib.check(L.pq_xchg_connect(h, arr), "connect")
    D, I = torch.zeros((100_000, 10), device="cuda"), torch.zeros((100_000, 10), dtype=torch.int64, device="cuda")   # 24 MB of results; buffers are 2 MB
    rc = L.pq_xchg_run(h, 0, 1, 100_000, 10, ctypes.c_void_p(D.data_ptr()), ctypes.c_void_p(I.data_ptr()), ctypes.c_void_p(D.data_ptr()),
                       ctypes.c_void_p(I.data_ptr()), 1, None)
    assert rc != 0 and "too small" in _lib.last_error()
    L.pq_xchg_free(h)
