"""-m gpu: pq_merge_shard_results through the C ABI — the kernel that follows the NCCL all-gather of the per-shard (D, I)
lists (proqa_b200/sharded.py) — against numpy, including the sizes beyond the in-CTA sort (two shards at k = 10000 is the
retrieval/trec_process.py:76 shape on two GPUs; 8 x 5000 the qa/online_sampler.py:113 shape on eight)."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
FLT_MAX = np.float32(np.finfo(np.float32).max)


def _lists(G, nq, k, metric, seed, coarse):
    rng = np.random.default_rng(seed)
    per = max(5000, 2 * k)
    D_all = np.empty((G, nq, k), np.float32)
    I_all = np.empty((G, nq, k), np.int64)
    for g in range(G):   # shard g owns ids [g*per, (g+1)*per); best-first lists, ties in ascending id order, some short lists
        for q in range(nq):
            n_valid = k if (g + q) % 3 else k // 2
            ids = np.sort(rng.choice(per, n_valid, replace=False)) + g * per
            sc = rng.standard_normal(n_valid).astype(np.float32)
            if coarse:
                sc = np.round(sc, 1) + np.float32(0.0)       # (+0.0: the kernels order by bit pattern, -0.0 below +0.0; the engine never emits -0.0)
            if metric == 1:
                sc = np.abs(sc)
            order = np.lexsort((ids, -sc if metric == 0 else sc))
            D_all[g, q, :n_valid], I_all[g, q, :n_valid] = sc[order], ids[order]
            D_all[g, q, n_valid:] = FLT_MAX if metric == 1 else -FLT_MAX
            I_all[g, q, n_valid:] = -1
    return D_all, I_all


def _reference(D_all, I_all, k, metric):
    G, nq, _ = D_all.shape
    D = D_all.transpose(1, 0, 2).reshape(nq, -1)
    I = I_all.transpose(1, 0, 2).reshape(nq, -1)
    Do = np.empty((nq, k), np.float32)
    Io = np.empty((nq, k), np.int64)
    for q in range(nq):
        valid = I[q] >= 0
        order = np.lexsort((I[q], -D[q] if metric == 0 else D[q], ~valid))[:k]
        Do[q], Io[q] = D[q, order], I[q, order]
        pad = ~valid[order]
        Do[q, pad] = FLT_MAX if metric == 1 else -FLT_MAX
        Io[q, pad] = -1
    return Do, Io


def _merge_on_gpu(D_all, I_all, k, metric, async_stream=False):
    import torch
    from proqa_b200 import _lib
    G, nq, _ = D_all.shape
    Dd, Id = torch.from_numpy(D_all).cuda(), torch.from_numpy(I_all).cuda()
    Do = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    Io = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    args = [0, metric, G, nq, k, ctypes.c_void_p(Dd.data_ptr()), ctypes.c_void_p(Id.data_ptr()), ctypes.c_void_p(Do.data_ptr()),
            ctypes.c_void_p(Io.data_ptr())]
    if async_stream:
        rc = _lib.lib().pq_merge_shard_results_async(*args, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    else:
        rc = _lib.lib().pq_merge_shard_results(*args)
    _lib.check(rc, "merge_shard_results")
    torch.cuda.synchronize()
    return Do.cpu().numpy(), Io.cpu().numpy()


@pytest.mark.parametrize("G,nq,k,metric,coarse", [(2, 7, 100, 0, True), (8, 5, 80, 1, True), (8, 3, 1000, 0, False),      # in-CTA sort
                                                  (2, 6, 10000, 0, False), (2, 3, 10000, 1, True), (8, 4, 5000, 0, True),  # by ranking
                                                  (3, 2, 15360, 1, False)])
def test_merge_shard_results_matches_numpy(G, nq, k, metric, coarse):
    D_all, I_all = _lists(G, nq, k, metric, 100 * G + k, coarse)
    D, I = _merge_on_gpu(D_all, I_all, k, metric, async_stream=(G == 8))
    Dr, Ir = _reference(D_all, I_all, k, metric)
    np.testing.assert_array_equal(I, Ir)
    np.testing.assert_array_equal(D.view(np.uint32), Dr.view(np.uint32))


def test_merge_l2_equal_distances_with_unsorted_ids():
    """Equal D inside one list with ids not ascending (possible for L2 after clamping/rounding): ties go by (list, position),
    every output slot is written exactly once — by both kernels."""
    rng = np.random.default_rng(3)
    for G, k in ((2, 64), (2, 10000)):
        nq = 3
        D_all = np.sort(np.round(np.abs(rng.standard_normal((G, nq, k))), 1).astype(np.float32), axis=2)
        I_all = np.stack([np.stack([rng.permutation(k) + g * k for _ in range(nq)]) for g in range(G)]).astype(np.int64)
        D, I = _merge_on_gpu(D_all, I_all, k, 1)
        Dc, Ic = D_all.transpose(1, 0, 2).reshape(nq, -1), I_all.transpose(1, 0, 2).reshape(nq, -1)
        for q in range(nq):
            order = np.argsort(Dc[q], kind="stable")[:k]
            np.testing.assert_array_equal(I[q], Ic[q, order])
            np.testing.assert_array_equal(D[q], Dc[q, order])


def test_merge_refuses_more_than_64_lists_beyond_the_sort():
    import torch
    from proqa_b200 import _lib
    G, k = 65, 300
    Dd = torch.zeros((G, 1, k), device="cuda")
    Id = torch.zeros((G, 1, k), dtype=torch.int64, device="cuda")
    Do, Io = torch.empty((1, k), device="cuda"), torch.empty((1, k), dtype=torch.int64, device="cuda")
    rc = _lib.lib().pq_merge_shard_results(0, 0, G, 1, k, ctypes.c_void_p(Dd.data_ptr()), ctypes.c_void_p(Id.data_ptr()),
                                           ctypes.c_void_p(Do.data_ptr()), ctypes.c_void_p(Io.data_ptr()))
    assert rc != 0 and "64 lists" in _lib.last_error()
