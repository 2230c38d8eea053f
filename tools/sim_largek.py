"""CPU simulation of the large-k tensor-tier plan: statistics only, no kernels.

For a synthetic corpus (iid Gaussian rows, or topic clusters stored in document order) and a handful of queries it computes
every row's bf16-filter score and exact fp32 score, then replays the three phases on the score arrays:

  A. thresholds from a sample (every `step`-th row, or every `step`-th tile of 128 rows): T = (k'-th best bf16 score of the sample) - 2E
  B. one pass over all rows admitting bf16 score >= T  -> survivors per query, spread over the candidate slabs
  C. finalize: A_k = k-th best bf16 score of the survivors, certificate T <= A_k - 2E, rows to rescore = survivors with
     bf16 score >= A_k - 2E, top-k by exact score compared with the brute-force exact top-k

and reports survivors / pool capacity / slab occupancy / certificate / recall.  Run: python tools/sim_largek.py [--n 8800000]
"""
import argparse
import math

import numpy as np


def bf16_round(x):
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def plan(k, n_rows):
    """Mirror of the host plan in pq_plan.h: plan_large_k."""
    rank_target = 1.5 * k
    step = max(2, math.ceil(rank_target / 1024))
    k_s = math.ceil(rank_target / step)
    pool = max(1 << (k - 1).bit_length(), min(27648, 2 * k))   # keys the finalize kernel keeps in shared memory (rescored set)
    return dict(step=step, k_sample=k_s, pool=pool)


def corpus_chunks(n, kind, seed, chunk=1 << 18):
    rng = np.random.default_rng(seed)
    centers = rng.standard_normal((4096, 128)).astype(np.float32) if kind == "topics" else None
    done = 0
    while done < n:
        m = min(chunk, n - done)
        if kind == "iid":
            x = rng.standard_normal((m, 128)).astype(np.float32)
        else:  # runs of 10..300 consecutive rows around one centre: articles cut into paragraphs, document order
            x = np.empty((m, 128), np.float32)
            i = 0
            while i < m:
                run = int(rng.integers(10, 300))
                c = centers[int(rng.integers(0, len(centers)))]
                j = min(m, i + run)
                x[i:j] = c + 0.7 * rng.standard_normal((j - i, 128)).astype(np.float32)
                i = j
        yield done, x
        done += m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8_800_000)
    ap.add_argument("--k", type=int, default=10000)
    ap.add_argument("--nq", type=int, default=16)
    ap.add_argument("--kind", default="iid", choices=["iid", "topics"])
    ap.add_argument("--slabs", type=int, default=16)
    ap.add_argument("--sample", default="rows", choices=["rows", "tiles"])
    args = ap.parse_args()
    n, k, nq = args.n, args.k, args.nq
    rngq = np.random.default_rng(4321)
    xq = rngq.standard_normal((nq, 128)).astype(np.float32)
    if args.kind == "topics":   # queries near topics, so that whole runs of rows score high together
        cs = np.random.default_rng(1234).standard_normal((4096, 128)).astype(np.float32)
        xq = cs[rngq.integers(0, 4096, nq)] + 0.5 * xq
    xq_b = bf16_round(xq)
    S = np.empty((nq, n), np.float32)   # exact fp32 (numpy matmul; the engine's own rounding differs by ~1e-6 relative)
    B = np.empty((nq, n), np.float32)   # bf16-filter score
    max_norm2 = max_res2 = 0.0
    for off, x in corpus_chunks(n, args.kind, 1234):
        xb = bf16_round(x)
        S[:, off:off + len(x)] = xq @ x.T
        B[:, off:off + len(x)] = xq_b @ xb.T
        max_norm2 = max(max_norm2, float((x.astype(np.float64) ** 2).sum(1).max()))
        max_res2 = max(max_res2, float(((x - xb).astype(np.float64) ** 2).sum(1).max()))
    p = plan(k, n)
    step, k_s, pool = p["step"], p["k_sample"], p["pool"]
    print(f"N={n} k={k} kind={args.kind}: step={step} k_sample={k_s} pool={pool}")
    samp = np.zeros(n, bool)
    if args.sample == "rows":       # every step-th row (gathered into a compact sample copy)
        samp[step // 2::step] = True
    else:                           # every step-th tile of 128 rows
        for t in range(0, (n + 127) // 128, step):
            samp[t * 128:(t + 1) * 128] = True
    C, rc = math.sqrt(max_norm2), math.sqrt(max_res2)
    rows = []
    for q in range(nq):
        Q = float(np.linalg.norm(xq[q].astype(np.float64)))
        rq = float(np.linalg.norm((xq[q] - xq_b[q]).astype(np.float64)))
        E = 1.0002 * (rq * (C + rc) + Q * rc) + 6.2e-5 * (Q + rq) * (C + rc) + 2.4e-6 * Q * C   # pq_mma_init_state_kernel
        assert np.abs(B[q] - S[q]).max() <= E
        bs = B[q][samp]
        a_s = np.partition(bs, len(bs) - k_s)[len(bs) - k_s]
        T = a_s - 2 * E
        surv = np.nonzero(B[q] >= T)[0]
        n_s = len(surv)
        slab_of = (surv // 128) % args.slabs      # tiles dealt round-robin to the slabs
        occ = np.bincount(slab_of, minlength=args.slabs).max()
        a_k = np.partition(B[q][surv], n_s - k)[n_s - k] if n_s >= k else np.nan
        resc = surv[B[q][surv] >= a_k - 2 * E] if n_s >= k else surv
        cert = n_s >= k and len(resc) <= pool and T <= a_k - 2 * E
        top = resc[np.argsort(-S[q][resc], kind="stable")[:k]]
        truth = np.argsort(-S[q], kind="stable")[:k]
        exact = cert and set(top.tolist()) == set(truth.tolist())
        rank_T = int((B[q] >= a_s).sum())
        rows.append((n_s, occ, len(resc), rank_T, cert, exact))
        print(f"q{q:2d}: 2E/score_k={2 * E / abs(a_k):.4f} rank(sample k')={rank_T:6d} survivors={n_s:6d} ({n_s / k:.2f}k) "
              f"max slab={occ:5d} rescored={len(resc):6d} certificate={'ok' if cert else 'FAIL'} exact_topk={'ok' if exact else '-'}")
    r = np.array([(a, b, c, d) for a, b, c, d, _, _ in rows], float)
    print(f"survivors/k: mean {r[:, 0].mean() / k:.2f} max {r[:, 0].max() / k:.2f}; rank estimate / k: min {r[:, 3].min() / k:.2f} "
          f"max {r[:, 3].max() / k:.2f}; certificates ok: {sum(x[4] for x in rows)}/{nq}; exact: {sum(x[5] for x in rows)}/{nq}")


if __name__ == "__main__":
    main()
