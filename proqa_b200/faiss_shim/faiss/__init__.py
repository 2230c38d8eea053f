"""``import faiss`` → the B200 engine.  Put ``proqa_b200/faiss_shim`` first on PYTHONPATH and the reference scripts
(retrieval/eval_retrieval.py, retrieval/group_paras.py, retrieval/trec_process.py) run unmodified.

Only the slice of the FAISS 1.6.3 Python API those scripts call is provided:

    faiss.IndexFlatIP(d), faiss.IndexFlatL2(d)       eval_retrieval.py:102, group_paras.py:36,38, trec_process.py:74
    index.add / search / reset / ntotal / d / is_trained / train
    faiss.Clustering(d, k) + .verbose/.niter/.max_points_per_centroid/.train(x, index)/.centroids   group_paras.py:40-46
    faiss.vector_float_to_array(v)                   group_paras.py:46

Importing this module does not load the native library and never touches CUDA: eval_retrieval.py forks its worker
pool (:92-96) after ``import faiss`` (:4); the device is initialised by the first ``add``/``search``.
There is no CPU fallback: without libproqa_b200.so or without a B200 those calls raise.
"""
import os as _os
import sys as _sys

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
if _ROOT not in _sys.path:
    _sys.path.insert(0, _ROOT)

from proqa_b200 import index as _index  # noqa: E402
from proqa_b200.index import METRIC_INNER_PRODUCT, METRIC_L2, IndexFlat  # noqa: E402,F401
from proqa_b200.multi import MultiGpuIndexFlat, make_index as _make_index  # noqa: E402,F401


class IndexFlatIP(_index.IndexFlatIP):
    """faiss.IndexFlatIP(d): one GPU — or, with PROQA_B200_DEVICES=0,1,... (or "all"), every named GPU of the box behind the
    same object (proqa_b200/multi.py), so that the unmodified scripts scale without a launcher."""
    def __new__(cls, d, *args):
        if not args and cls is IndexFlatIP:
            ix = _make_index(d, METRIC_INNER_PRODUCT)
            if isinstance(ix, MultiGpuIndexFlat):
                return ix
        return super().__new__(cls)


class IndexFlatL2(_index.IndexFlatL2):
    def __new__(cls, d, *args):
        if not args and cls is IndexFlatL2:
            ix = _make_index(d, METRIC_L2)
            if isinstance(ix, MultiGpuIndexFlat):
                return ix
        return super().__new__(cls)
from proqa_b200.clustering import Clustering, ClusteringParameters, vector_float_to_array, vector_to_array  # noqa: E402,F401

__version__ = "1.6.3+proqa_b200"
