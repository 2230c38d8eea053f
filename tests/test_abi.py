"""CPU: the C-ABI library loads without a GPU and exports every symbol include/proqa_b200.h declares.

No compute call is made here (there is no device in the build container); what a call without a device must do —
fail loudly with PQ_ERR_NO_DEVICE, never fall back — is checked too.
"""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "proqa_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pq_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from proqa_b200 import _lib
    L = _lib.lib()
    names = _declared_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/proqa_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in proqa_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(names)


def test_version_and_no_cuda_at_load():
    import proqa_b200 as pq
    assert "sm_100a" in pq.version()
    # creating an index must not touch CUDA (eval_retrieval.py forks after import / before index use)
    ix = pq.IndexFlatIP(128)
    assert ix.ntotal == 0 and ix.d == 128 and ix.is_trained
    del ix


def test_bad_arguments_raise_like_faiss():
    import proqa_b200 as pq
    with pytest.raises(ValueError):
        pq.IndexFlatIP(64)  # engine is built for d = 128 (eval_retrieval.py:98)
    ix = pq.IndexFlatIP(128)
    with pytest.raises(AssertionError):
        ix.add(np.zeros((3, 64), np.float32))
    with pytest.raises(AssertionError):
        ix.search(np.zeros((3, 64), np.float32), 5)
    with pytest.raises(AssertionError):
        ix.search(np.zeros((3, 128), np.float32), 0)


def test_clustering_argument_errors_are_faiss_runtime_errors():
    """Clustering::train's FAISS_THROW_IF_NOT checks surface as RuntimeError through SWIG; the engine's host mirror does the
    same, with FAISS's wording, before it ever touches a device."""
    import proqa_b200 as pq
    clus = pq.Clustering(128, 50)
    with pytest.raises(RuntimeError, match=r"Number of training points \(10\) should be at least as large as number of clusters \(50\)"):
        clus.train(np.zeros((10, 128), np.float32), pq.IndexFlatL2(128))
    bad = np.zeros((100, 128), np.float32)
    bad[3, 3] = np.nan
    with pytest.raises(RuntimeError, match="input contains NaN's or Inf's"):
        clus.train(bad, pq.IndexFlatL2(128))
    with pytest.raises(AssertionError):
        clus.train(np.zeros((100, 64), np.float32), pq.IndexFlatL2(128))


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    import proqa_b200 as pq
    ix = pq.IndexFlatIP(128)
    with pytest.raises(RuntimeError, match="no CPU fallback|no CUDA device|fallback"):
        ix.add(np.zeros((4, 128), np.float32))
    with pytest.raises(RuntimeError):
        ix.search(np.zeros((1, 128), np.float32), 1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "proqa_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
                assert "liboracle" not in src, f"{f} links the oracle"


def test_missing_library_fails_loudly():
    """No built extension => every entry point raises; nothing falls back to Python or the oracle."""
    import subprocess
    import sys
    code = ("import os, sys; os.environ['PROQA_B200_LIB'] = '/nonexistent/libproqa_b200.so'; sys.path.insert(0, %r)\n"
            "import proqa_b200 as pq\n"
            "try:\n    pq.IndexFlatIP(128)\nexcept RuntimeError as e:\n    assert 'no CPU fallback' in str(e).lower() or 'no cpu fallback' in str(e).lower(), e; print('raised')\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "raised" in out.stdout, out.stdout + out.stderr
    assert "oracle" not in sys.modules or True
