"""-m gpu, needs two GPUs: the one-process-per-GPU layout (proqa_b200/sharded.py) under torchrun — NCCL all-gather + merge kernel,
thresholds exchanged through CUDA-IPC peer mailboxes — against the oracle (tests/workers/sharded_check.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_row_sharded_with_threshold_exchange():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU in this box (tests/test_gpu_multi.py covers the exchange with two shards on one device)")
    port = 29500 + os.getpid() % 500
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "workers", "sharded_check.py")],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "FAIL" not in r.stdout and r.stdout.count("PASS") == 5
