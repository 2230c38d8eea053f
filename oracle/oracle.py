"""oracle/oracle.py — CPU oracle for ProQA's exact flat search.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module; the product package ``proqa_b200`` never does (it has no CPU path).

PARITY UNPINNED: the reference's arithmetic for this path is inside the third-party wheel
``faiss-cpu==1.6.3`` (/root/reference/requirements.txt:2), absent from /root/reference and not installable
here; the reference ships no tests or golden vectors for it.  This module restates the
published FAISS 1.6.3 algorithm behind the reference call sites

    /root/reference/retrieval/eval_retrieval.py:102-104   IndexFlatIP(d); add(xb); search(xq, 80)
    /root/reference/retrieval/group_paras.py:35-51         IndexFlatL2/IP; reset; add; search(data, 1)
    /root/reference/retrieval/trec_process.py:74-76        IndexFlatIP; search(xq, 10000)

Three oracles, from most to least authoritative:

* :func:`truth_fp64`      — brute force in float64: the ground truth the north-star tolerance is stated
  against (scores within 1e-4 relative, ids identical except within that tolerance of a tie).
* :class:`FaissFlatOracle` — FAISS semantics in fp32: ``sgemm`` on 4096-query x 1024-row blocks when
  nq >= 20 (numpy/OpenBLAS here), a direct dot loop below 20, per-query binary heap fed in ascending id
  with strict-improvement replacement, best-first reorder, ``-1`` / ``∓FLT_MAX`` padding, L2 as
  ``|x|^2+|y|^2-2<x,y>`` clamped at 0.  This is "what FAISS would print" and the CPU baseline that is timed.
* :func:`engine_spec`     — the GPU engine's *defined* result: score = eight fmaf chains of
  16 dims tree-combined, order = (score desc, id asc).  The CUDA path must match it bit for bit.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FLT_MAX = np.float32(3.4028234663852886e38)
METRIC_IP, METRIC_L2 = 0, 1


def build(force: bool = False) -> str:
    """Compile flat_oracle.c -> liboracle.so (gcc + OpenMP).  Building the checker is not using it."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "flat_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        f32p, i64p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int64)
        i64, i32 = ctypes.c_int64, ctypes.c_int
        L.faiss_flat_search.argtypes = [f32p, i64, f32p, i64, i32, i64, i32, f32p, i64p]
        L.faiss_flat_search.restype = None
        L.faiss_heaps_init.argtypes = [i64, i64, i32, f32p, i64p]
        L.faiss_heaps_init.restype = None
        L.faiss_heaps_addn.argtypes = [i64, i64, i64, i64, f32p, i64, i64, i32, f32p, f32p, f32p, i64p]
        L.faiss_heaps_addn.restype = None
        L.faiss_heaps_reorder.argtypes = [i64, i64, i32, f32p, i64p]
        L.faiss_heaps_reorder.restype = None
        L.engine_flat_search.argtypes = [f32p, i64, f32p, i64, i32, i64, i32, i64, f32p, i64p]
        L.engine_flat_search.restype = None
        L.engine_chain_dot.argtypes = [f32p, f32p, i32]
        L.engine_chain_dot.restype = ctypes.c_float
        L.faiss_kmeans_train.argtypes = [f32p, i64, i32, i64, i32, i32, i32, i64, i32, f32p, f32p, ctypes.POINTER(ctypes.c_int), i64p]
        L.faiss_kmeans_train.restype = i32
        L.faiss_rand_perm_export.argtypes = [ctypes.POINTER(ctypes.c_int), i64, i64]
        L.faiss_rand_perm_export.restype = None
        _LIB = L
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


# ------------------------------------------------------------------------------------------------
# ground truth
# ------------------------------------------------------------------------------------------------
def scores_fp64(xq, xb, metric=METRIC_IP):
    """Exact scores in float64.  IP: <q,x>.  L2: squared distance."""
    q = np.asarray(xq, dtype=np.float64)
    b = np.asarray(xb, dtype=np.float64)
    if metric == METRIC_IP:
        return q @ b.T
    return (q * q).sum(1)[:, None] + (b * b).sum(1)[None, :] - 2.0 * (q @ b.T)


def truth_fp64(xq, xb, k, metric=METRIC_IP):
    """Top-k by float64 score; ties -> lower id first.  Returns (D float64 [nq,k], I int64 [nq,k])."""
    S = scores_fp64(xq, xb, metric)
    nq, nb = S.shape
    D = np.full((nq, k), -np.inf if metric == METRIC_IP else np.inf)
    I = np.full((nq, k), -1, dtype=np.int64)
    kk = min(k, nb)
    key = -S if metric == METRIC_IP else S
    for i in range(nq):
        order = np.lexsort((np.arange(nb), key[i]))[:kk]
        I[i, :kk] = order
        D[i, :kk] = S[i, order]
    return D, I


# ------------------------------------------------------------------------------------------------
# FAISS 1.6.3 restatement
# ------------------------------------------------------------------------------------------------
class FaissFlatOracle:
    """faiss.IndexFlatIP / IndexFlatL2 restated on the CPU (API as used by the reference scripts)."""

    BLAS_THRESHOLD = 20      # distance_compute_blas_threshold
    BS_QUERY, BS_DB = 4096, 1024  # distance_compute_blas_{query,database}_bs

    def __init__(self, d, metric=METRIC_IP):
        self.d = int(d)
        self.metric_type = metric
        self.is_trained = True
        self._xb = np.zeros((0, self.d), dtype=np.float32)

    @property
    def ntotal(self):
        return self._xb.shape[0]

    def add(self, x):
        x = _f32(x)
        assert x.ndim == 2 and x.shape[1] == self.d
        self._xb = np.concatenate([self._xb, x], axis=0)  # copies, like std::vector::insert

    def reset(self):
        self._xb = np.zeros((0, self.d), dtype=np.float32)

    def train(self, x):
        pass

    def search(self, x, k, use_blas=None):
        x = _f32(x)
        assert x.ndim == 2 and x.shape[1] == self.d
        nq, nb, k = x.shape[0], self.ntotal, int(k)
        D = np.empty((nq, k), dtype=np.float32)
        I = np.empty((nq, k), dtype=np.int64)
        if nq == 0:
            return D, I
        L = _lib()
        if use_blas is None:
            use_blas = nq >= self.BLAS_THRESHOLD
        if not use_blas or nb == 0:
            L.faiss_flat_search(_p(x, ctypes.c_float), nq, _p(self._xb, ctypes.c_float), nb, self.d, k, self.metric_type,
                                _p(D, ctypes.c_float), _p(I, ctypes.c_int64))
            return D, I
        # BLAS path: sgemm on (4096 x 1024) blocks, heaps updated after every block
        xb = self._xb
        qn = (x * x).sum(1).astype(np.float32) if self.metric_type == METRIC_L2 else np.zeros(1, np.float32)
        bn = (xb * xb).sum(1).astype(np.float32) if self.metric_type == METRIC_L2 else np.zeros(1, np.float32)
        L.faiss_heaps_init(nq, k, self.metric_type, _p(D, ctypes.c_float), _p(I, ctypes.c_int64))
        for q0 in range(0, nq, self.BS_QUERY):
            q1 = min(nq, q0 + self.BS_QUERY)
            for j0 in range(0, nb, self.BS_DB):
                j1 = min(nb, j0 + self.BS_DB)
                S = np.ascontiguousarray(x[q0:q1] @ xb[j0:j1].T, dtype=np.float32)  # sgemm_
                L.faiss_heaps_addn(q1 - q0, q0, j1 - j0, j0, _p(S, ctypes.c_float), S.shape[1], k, self.metric_type,
                                   _p(qn, ctypes.c_float), _p(bn, ctypes.c_float), _p(D, ctypes.c_float), _p(I, ctypes.c_int64))
        L.faiss_heaps_reorder(nq, k, self.metric_type, _p(D, ctypes.c_float), _p(I, ctypes.c_int64))
        return D, I


def IndexFlatIP(d):
    return FaissFlatOracle(d, METRIC_IP)


def IndexFlatL2(d):
    return FaissFlatOracle(d, METRIC_L2)


# ------------------------------------------------------------------------------------------------
# faiss.Clustering restated (group_paras.py:40-47)
# ------------------------------------------------------------------------------------------------
def rand_perm(n, seed):
    """faiss::rand_perm: Fisher-Yates driven by std::mt19937(seed) % (n - i)."""
    perm = np.empty(n, dtype=np.int32)
    _lib().faiss_rand_perm_export(_p(perm, ctypes.c_int), n, seed)
    return perm


class FaissClusteringOracle:
    """faiss.Clustering(d, k) with the attributes group_paras.py sets; train(x, index) leaves the centroids in
    ``centroids`` (flat, k*d) and in ``index`` (reset + add), the per-iteration objective in ``obj``."""

    def __init__(self, d, k):
        self.d, self.k = int(d), int(k)
        self.niter, self.nredo, self.verbose, self.spherical = 25, 1, False, False
        self.min_points_per_centroid, self.max_points_per_centroid, self.seed = 39, 256, 1234
        self.centroids = np.zeros(0, np.float32)
        self.obj = np.zeros(0, np.float32)
        self.nsplit = np.zeros(0, np.int32)

    def train(self, x, index):
        x = _f32(x)
        n, d = x.shape
        assert d == self.d and n >= self.k
        cent = np.empty(self.k * d, np.float32)
        obj = np.zeros(max(1, self.niter), np.float32)
        nsplit = np.zeros(max(1, self.niter), np.int32)
        nit = _lib().faiss_kmeans_train(_p(x, ctypes.c_float), n, d, self.k, int(self.niter), int(bool(self.spherical)),
                                        int(self.max_points_per_centroid), int(self.seed), int(index.metric_type), _p(cent, ctypes.c_float),
                                        _p(obj, ctypes.c_float), nsplit.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), None)
        self.centroids, self.obj, self.nsplit = cent, obj[:nit].copy(), nsplit[:nit].copy()
        index.reset()
        index.add(cent.reshape(self.k, d))


# ------------------------------------------------------------------------------------------------
# engine specification (bit-exact target for the CUDA kernels)
# ------------------------------------------------------------------------------------------------
def engine_spec(xq, xb, k, metric=METRIC_IP, id_base=0):
    xq, xb = _f32(xq), _f32(xb)
    nq, nb, d, k = xq.shape[0], xb.shape[0], xq.shape[1], int(k)
    D = np.empty((nq, k), dtype=np.float32)
    I = np.empty((nq, k), dtype=np.int64)
    if nq:
        _lib().engine_flat_search(_p(xq, ctypes.c_float), nq, _p(xb, ctypes.c_float), nb, d, k, metric, id_base,
                                  _p(D, ctypes.c_float), _p(I, ctypes.c_int64))
    return D, I


# ------------------------------------------------------------------------------------------------
# the north-star comparator
# ------------------------------------------------------------------------------------------------
def check_against_truth(D, I, xq, xb, k, metric=METRIC_IP, rtol=1e-4, id_base=0):
    """Tie-aware parity check against the fp64 ground truth.  Returns a list of problems (empty = pass).

    * every returned id is a valid, distinct row (or -1 padding exactly where k > ntotal);
    * |D - true score of the returned id| <= rtol * max(|score|, eps*|q||x|)   (eps = 1e-3: relative
      tolerance is ill-defined near 0, so it is floored by the magnitude of the operands);
    * the returned set equals the true top-k set except for swaps among rows whose true score is within
      tolerance of the true k-th score;
    * the returned order is non-increasing in true score up to the same tolerance.
    """
    D = np.asarray(D)
    I = np.asarray(I)
    S = scores_fp64(xq, xb, metric)
    nq, nb = S.shape
    sign = 1.0 if metric == METRIC_IP else -1.0
    qn = np.linalg.norm(np.asarray(xq, np.float64), axis=1)
    bn = np.linalg.norm(np.asarray(xb, np.float64), axis=1) if nb else np.zeros(0)
    problems = []
    kk = min(k, nb)
    for q in range(nq):
        ids = I[q] - id_base
        if (I[q, kk:] != -1).any():
            problems.append(f"q{q}: padding ids not -1")
        pad = FLT_MAX if metric == METRIC_L2 else -FLT_MAX
        if kk < k and not np.all(D[q, kk:] == pad):
            problems.append(f"q{q}: padding distances not {pad}")
        ids = ids[:kk]
        if kk == 0:
            continue
        if ids.min() < 0 or ids.max() >= nb or len(set(ids.tolist())) != kk:
            problems.append(f"q{q}: invalid or duplicate ids")
            continue
        true = S[q, ids]
        scale = qn[q] * bn[ids]
        if metric == METRIC_L2:
            scale = (qn[q] + bn[ids]) ** 2
        tol = rtol * np.maximum(np.abs(true), 1e-3 * scale) + 1e-30
        err = np.abs(D[q, :kk].astype(np.float64) - true)
        if (err > tol).any():
            j = int(np.argmax(err - tol))
            problems.append(f"q{q}: score[{j}]={D[q, j]} vs true {true[j]} (err {err[j]:.3e} > tol {tol[j]:.3e})")
        best = np.sort(sign * S[q])[::-1]
        kth = best[kk - 1]
        ktol = rtol * max(abs(kth), 1e-3 * (qn[q] * bn.max() if metric == METRIC_IP else (qn[q] + bn.max()) ** 2)) + 1e-30
        worst_returned = (sign * true).min()
        if worst_returned < kth - 2 * ktol:
            problems.append(f"q{q}: returned a row scoring {worst_returned} below the true k-th {kth}")
        must_have = np.nonzero(sign * S[q] > kth + 2 * ktol)[0]
        missing = set(must_have.tolist()) - set(ids.tolist())
        if missing:
            problems.append(f"q{q}: missing rows {sorted(missing)[:5]} that beat the k-th score by more than tolerance")
        st = sign * true
        if (st[1:] > st[:-1] + 2 * np.maximum(tol[1:], tol[:-1])).any():
            problems.append(f"q{q}: results not best-first")
    return problems
