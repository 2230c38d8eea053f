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
