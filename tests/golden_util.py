"""Helpers around tests/golden/eval_fixture.npz (made by tests/golden/make_fixtures.py from the reference's own
retrieval/eval_retrieval.py).  Pure integer work: nothing here reads /root/reference."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def load_eval_fixture():
    z = np.load(os.path.join(HERE, "golden", "eval_fixture.npz"))
    xb = z["xb"].astype("float32")   # eval_retrieval.py:100  np.load(indexpath).astype('float32')
    xq = z["xq"].astype("float32")   # eval_retrieval.py:99
    has = np.unpackbits(z["has_answer"], axis=1)[:, :xb.shape[0]].astype(bool)
    return dict(xb=xb, xq=xq, I=z["I"].astype(np.int64), D=z["D"], has_answer=has, topk=int(z["topk"]),
                recall_lines=[str(s) for s in z["recall_lines"]])


def recall_lines(I, fx):
    """eval_retrieval.py:47-65,115-123 restated on the committed has-answer matrix: same cut-offs, same key order,
    same format string, np.mean of the 0/1 hits."""
    topk = fx["topk"]
    has = fx["has_answer"]
    nq = I.shape[0]
    covered = np.stack([has[q, I[q, :topk]] for q in range(nq)]).astype(np.int64)
    agg = {str(topk): [int(np.sum(c) > 0) for c in covered]}
    for cut in (5, 10, 20, 50):
        agg[str(cut)] = [int(np.sum(c[:cut]) > 0) for c in covered]
    return ['Top {} Recall for {} QA pairs: {} ...'.format(k, len(v), np.mean(v)) for k, v in agg.items()]


def load_trec_fixture():
    """tests/golden/trec_fixture.npz (retrieval/trec_process.py: retrieve_topk, k = 10000).  The inputs are regenerated from
    the committed seed and checked against the committed digests."""
    import hashlib
    import json
    z = np.load(os.path.join(HERE, "golden", "trec_fixture.npz"))
    rng = np.random.default_rng(int(z["seed"]))
    xb = rng.standard_normal((int(z["n"]), 128)).astype(np.float16)
    xq = rng.standard_normal((int(z["nq"]), 128)).astype(np.float16)
    assert hashlib.sha1(xb.tobytes()).hexdigest() == str(z["xb_sha1"]) and hashlib.sha1(xq.tobytes()).hexdigest() == str(z["xq_sha1"]), \
        "numpy generated different inputs than when the fixture was made"
    return dict(xb=xb.astype("float32"), xq=xq.astype("float32"), k=int(z["k"]), I=z["I"].astype(np.int64),
                labels=[json.loads(str(s)) for s in z["labels"]], recall_line=str(z["recall_line"]))


def trec_recall_line(I, fx):
    """trec_process.py:82-91 restated: a query is covered when any of its labels is among the ids returned."""
    covered = [int(np.sum([int(int(j) in fx["labels"][q]) for j in I[q]]) > 0) for q in range(I.shape[0])]
    return f"Avg recall: {np.mean(covered)}"


def assert_same_up_to_near_ties(I, I_ref, xq, xb, rtol=1e-4):
    """Position by position the two results name the same row, or two rows whose fp64 scores agree within the north-star
    tolerance (a near-tie that fp32 accumulation order may resolve either way)."""
    S = xq.astype(np.float64) @ xb.astype(np.float64).T
    a = np.take_along_axis(S, I, 1)
    b = np.take_along_axis(S, I_ref, 1)
    assert (np.abs(a - b) <= rtol * np.maximum(np.abs(b), 1e-3) + 1e-6).all()
    assert (I == I_ref).mean() > 0.98
