"""TEST INFRASTRUCTURE: a module named ``faiss`` backed by the CPU oracle (oracle/oracle.py).

Put this directory first on PYTHONPATH to run the reference's scripts (retrieval/eval_retrieval.py:4 does
``import faiss``) unmodified on the FAISS-1.6.3 restatement.  tests/golden/make_fixtures.py uses it to produce the
committed golden vectors; nothing in the product imports it.  With PROQA_GOLDEN_DUMP=<path.npz> every search() result
is also written out, because eval_retrieval.py never stores I.
"""
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from oracle import oracle as _o  # noqa: E402

METRIC_INNER_PRODUCT, METRIC_L2 = 0, 1


class _Dumping(_o.FaissFlatOracle):
    def search(self, x, k):
        D, I = super().search(x, k)
        path = os.environ.get("PROQA_GOLDEN_DUMP")
        if path:
            np.savez(path, D=D, I=I)
        return D, I


def IndexFlatIP(d):
    return _Dumping(d, _o.METRIC_IP)


def IndexFlatL2(d):
    return _Dumping(d, _o.METRIC_L2)


Clustering = _o.FaissClusteringOracle


def vector_float_to_array(v):
    return np.array(v, dtype=np.float32, copy=True)
