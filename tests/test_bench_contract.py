"""CPU: bench.py's reference arm (the FAISS restatement on host cores) prints ONE JSON line with the contract's keys.
The GPU arm cannot run here (no device): it must refuse loudly instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e, cwd=ROOT, timeout=600)


def test_reference_arm_json_line():
    out = _run("--impl", "reference", "--workload", "c1", "--rows", "60000", "--nq", "64", "--steps", "1", "--warmup", "0")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "queries_per_sec" and d["unit"] == "queries/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["vs_baseline"] is None and d["higher_is_better"] is True and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and d["value"] > 0


def test_reference_arm_other_ranks_stay_silent():
    out = _run("--impl", "reference", "--workload", "c1", "--rows", "20000", "--nq", "8", "--steps", "1", "--warmup", "0", "--gpus", "2",
               env={"RANK": "1", "WORLD_SIZE": "2"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_refuses_without_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a device is present")
    out = _run("--workload", "c1", "--rows", "20000", "--nq", "8", "--steps", "1", "--warmup", "0")
    assert out.returncode != 0
    assert "no CUDA device" in (out.stderr + out.stdout)
