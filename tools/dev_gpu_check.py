"""Developer diagnostics for a gpurun call (not a test, not a benchmark): parity + rough timings per tier."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import proqa_b200 as pq  # noqa: E402
from oracle import oracle  # noqa: E402


def run(tier, xb, xq, k, ref=None, metric=0, reps=3):
    ix = pq.IndexFlatIP(128) if metric == 0 else pq.IndexFlatL2(128)
    ix.set_tier(tier)
    t0 = time.time()
    ix.add(xb)
    t_add = time.time() - t0
    D, I = ix.search(xq, k)
    best = 1e9
    for _ in range(reps):
        t0 = time.time()
        D, I = ix.search(xq, k)
        best = min(best, time.time() - t0)
    st = ix.last_stats
    msg = f"[{tier}] nq={len(xq)} nb={len(xb)} k={k} add={t_add:.3f}s search={best*1e3:.2f}ms dev={st[6]/1e3:.2f}ms stats={st}"
    if ref is not None:
        Dr, Ir = ref
        ok_i = np.array_equal(I, Ir)
        ok_d = np.array_equal(D.view(np.uint32), Dr.view(np.uint32))
        msg += f" ids_exact={ok_i} scores_exact={ok_d}"
        if not ok_i:
            bad = np.nonzero((I != Ir).any(1))[0]
            msg += f" bad_queries={len(bad)} first={bad[:5].tolist()}"
            q = bad[0]
            msg += f"\n   got  I={I[q,:8].tolist()} D={D[q,:8].tolist()}\n   want I={Ir[q,:8].tolist()} D={Dr[q,:8].tolist()}"
    print(msg, flush=True)
    return D, I


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    rng = np.random.default_rng(1)
    if which in ("all", "small"):
        xb = rng.standard_normal((20000, 128), dtype=np.float32)
        xq = rng.standard_normal((130, 128), dtype=np.float32)
        ref = oracle.engine_spec(xq, xb, 80, 0)
        run("fp32", xb, xq[:16], 80, (ref[0][:16], ref[1][:16]))
        run("fp32", xb, xq, 80, ref)
        run("bf16", xb, xq, 80, ref)
    if which in ("all", "mid"):
        xb = rng.standard_normal((1000000, 128), dtype=np.float32)
        xq = rng.standard_normal((2032, 128), dtype=np.float32)
        ref = oracle.engine_spec(xq[:64], xb, 80, 0)
        run("fp32", xb, xq[:1], 80, (ref[0][:1], ref[1][:1]))
        run("fp32", xb, xq[:8], 80, (ref[0][:8], ref[1][:8]))
        run("fp32", xb, xq[:16], 80, (ref[0][:16], ref[1][:16]))
        run("fp32", xb, xq[:64], 80, ref)
        D, I = run("bf16", xb, xq, 80)
        print("   bf16 first-64 exact:", np.array_equal(I[:64], ref[1]), np.array_equal(D[:64].view(np.uint32), ref[0].view(np.uint32)), flush=True)
        run("bf16", xb, xq[:64], 80, ref)
        run("bf16", xb, xq[:256], 80)


if __name__ == "__main__":
    main()
