"""proqa_b200 — B200-native exact top-k maximum-inner-product search for ProQA's retrieval hot path.

Drop-in for the FAISS calls the reference makes (``faiss.IndexFlatIP(d)``, ``.add``, ``.search``,
``.reset`` — /root/reference/retrieval/eval_retrieval.py:102-104, retrieval/group_paras.py:35-51):
put ``proqa_b200/faiss_shim`` first on ``PYTHONPATH`` and the reference scripts run unmodified.
All compute happens in hand-written sm_100a CUDA kernels behind the C ABI in ``include/proqa_b200.h``;
there is no CPU fallback — without the built library or without a B200 the calls raise.
"""
from .index import (METRIC_INNER_PRODUCT, METRIC_L2, IndexFlat, IndexFlatIP, IndexFlatL2, last_search_stats)  # noqa: F401
from .clustering import Clustering, ClusteringParameters, vector_float_to_array  # noqa: F401
from ._lib import library_path, version  # noqa: F401
from .idmap import IdMap  # noqa: F401
from .multi import MultiGpuIndexFlat  # noqa: F401

__all__ = ["IndexFlat", "IndexFlatIP", "IndexFlatL2", "METRIC_INNER_PRODUCT", "METRIC_L2", "library_path", "version",
           "last_search_stats", "Clustering", "ClusteringParameters", "vector_float_to_array", "IdMap", "MultiGpuIndexFlat"]
