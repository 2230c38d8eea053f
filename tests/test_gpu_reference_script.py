"""-m gpu: the reference's UNMODIFIED retrieval/eval_retrieval.py on a B200 through the faiss shim — import faiss, fork the worker
pool (eval_retrieval.py:92-96) BEFORE the first CUDA call, IndexFlatIP(d), add, search (eval_retrieval.py:102-104), id mapping and
recall scoring — and the five recall lines it prints, byte for byte against the committed golden lines.

The script and its scratch tree are staged by tools/stage_reference_run.py (run in the build container, where /root/reference
exists) under baseline/_ref/ — git-ignored, but part of the snapshot the GPU box receives.  Skipped when it has not been staged."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUN = os.path.join(ROOT, "baseline", "_ref", "eval_run")


def _run(extra_env):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.path.join(ROOT, "proqa_b200", "faiss_shim") + os.pathsep + env.get("PYTHONPATH", "")
    env.update(extra_env)
    out = subprocess.run([sys.executable, "eval_retrieval.py", "../qa.jsonl", "../para_embed.npy", "../query_embed.npy", "../paras.db",
                          "--topk", "80", "--num-workers", "2"], cwd=os.path.join(RUN, "retrieval"), env=env, capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-3000:]
    return [ln for ln in out.stdout.splitlines() if ln.startswith("Top ")]


@pytest.mark.skipif(not os.path.exists(os.path.join(RUN, "retrieval", "eval_retrieval.py")), reason="reference script not staged (tools/stage_reference_run.py)")
@pytest.mark.parametrize("devices", [None, "0,0"])
def test_unmodified_eval_retrieval_prints_the_golden_recall_lines(devices):
    """devices=None: one GPU.  "0,0": the shim hands the script a MultiGpuIndexFlat (two shards; the 2000-row fixture is below the
    sharding threshold, so rows are replicated and the 24 queries split) — the same script, no launcher."""
    golden = [str(s) for s in np.load(os.path.join(ROOT, "tests", "golden", "eval_fixture.npz"))["recall_lines"]]
    lines = _run({} if devices is None else {"PROQA_B200_DEVICES": devices})
    assert lines == golden
