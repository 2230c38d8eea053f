"""CPU, build container only (needs /root/reference; skipped elsewhere): the reference's UNMODIFIED retrieval/eval_retrieval.py
run with ``PYTHONPATH=proqa_b200/faiss_shim`` imports the engine as ``faiss``, forks its worker pool, builds the index object
and reaches ``index.add`` — where, with no device in this container, the engine must refuse loudly (no CPU fallback)."""
import json
import os
import sqlite3
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = "/root/reference/retrieval/eval_retrieval.py"


@pytest.mark.skipif(not os.path.exists(SCRIPT), reason="reference tree not present (GPU box)")
def test_unmodified_eval_retrieval_binds_to_the_engine(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present: the GPU tests cover the search itself")
    (tmp_path / "retrieval").mkdir()
    (tmp_path / "pretrained_models").mkdir()
    np.save(tmp_path / "para.npy", np.random.default_rng(0).standard_normal((50, 128)).astype(np.float16))
    np.save(tmp_path / "q.npy", np.random.default_rng(1).standard_normal((3, 128)).astype(np.float16))
    json.dump({i: f"d{i}" for i in range(50)}, open(tmp_path / "pretrained_models" / "idx_id.json", "w"))
    con = sqlite3.connect(tmp_path / "p.db")
    con.execute("CREATE TABLE documents (id PRIMARY KEY, text)")
    con.executemany("INSERT INTO documents VALUES (?,?)", [(f"d{i}", "some text") for i in range(50)])
    con.commit()
    con.close()
    with open(tmp_path / "qa.jsonl", "w") as f:
        for _ in range(3):
            f.write(json.dumps({"question": "q ?", "answer": ["text"]}) + "\n")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.path.join(ROOT, "proqa_b200", "faiss_shim") + os.pathsep + env.get("PYTHONPATH", "")
    out = subprocess.run([sys.executable, SCRIPT, "../qa.jsonl", "../para.npy", "../q.npy", "../p.db", "--topk", "5", "--num-workers", "2"],
                         cwd=tmp_path / "retrieval", env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode != 0
    assert "proqa_b200: add failed" in out.stderr and "no CPU fallback" in out.stderr, out.stderr[-1500:]
