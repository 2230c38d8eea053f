#!/bin/bash
# Round 2, GPU call 1: everything written after round 1's last hardware run, each step under its own hard timeout.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/a_smi.log 2>&1
python - > $O/a_faiss_probe.log 2>&1 <<'PY'
# is a real FAISS reachable on the GPU box?
try:
    import faiss
    print("faiss importable:", faiss.__version__, faiss.__file__)
except Exception as e:
    print("faiss not importable:", repr(e))
import subprocess
print(subprocess.run("find / -iname '*faiss*' -not -path '*/proc/*' 2>/dev/null | grep -v graft | head -20", shell=True, capture_output=True, text=True).stdout)
PY
cat $O/a_faiss_probe.log
export PROQA_B200_LARGEK=1
timeout -s KILL 600 python -m pytest tests/test_gpu_largek.py -m gpu -x -q > $O/a_largek_tests.log 2>&1
echo "largek tests exit $?"; tail -25 $O/a_largek_tests.log
export PROQA_B200_STAGED_KMEANS=1
timeout -s KILL 600 python -m pytest tests/test_gpu_sharded_kmeans.py tests/test_gpu_merge.py -m gpu -q > $O/a_staged_merge_tests.log 2>&1
echo "staged k-means + merge tests exit $?"; tail -15 $O/a_staged_merge_tests.log
unset PROQA_B200_LARGEK
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > $O/a_pytest_all.log 2>&1
echo "all gpu tests exit $?"; tail -8 $O/a_pytest_all.log
for tool in memcheck racecheck; do
  PROQA_B200_SANITIZE_NEW=1 timeout -s KILL 600 compute-sanitizer --tool $tool --error-exitcode 3 python tools/sanitize_small.py > $O/a_new_paths_$tool.log 2>&1
  echo "$tool rc=$?"; tail -3 $O/a_new_paths_$tool.log
done
# large-k timing: trec shape (8.8M rows, k=10000), tensor tier vs scan
PROQA_B200_LARGEK=1 timeout -s KILL 300 python - > $O/a_largek_timing.log 2>&1 <<'PY'
import time, numpy as np, torch, proqa_b200 as pq
N, k = 8_800_000, 10000
g = torch.Generator(device="cuda"); g.manual_seed(1)
xb = torch.randn((N, 128), generator=g, device="cuda")
ix = pq.IndexFlatIP(128); ix.add_device(xb.data_ptr(), N); del xb
for nq in (16, 256, 1024):
    xq = np.random.default_rng(nq).standard_normal((nq, 128), dtype=np.float32)
    for rep in range(3):
        t0 = time.perf_counter(); D, I = ix.search(xq, k); dt = time.perf_counter() - t0
    print(f"largek nq={nq} k={k} rows={N}: {dt*1e3:.1f} ms  stats={ix.last_stats}", flush=True)
ix.set_tier("fp32")
xq = np.random.default_rng(0).standard_normal((4, 128), dtype=np.float32)
t0 = time.perf_counter(); D, I = ix.search(xq, k); dt = time.perf_counter() - t0
print(f"scan nq=4: {dt*1e3:.1f} ms")
PY
cat $O/a_largek_timing.log
