#!/bin/bash
# Round 2, first GPU call for the large-k tensor tier: parity, then sanitizer on a small case.
# Fail fast: every step under its own hard timeout so that a hang cannot hold the box.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PROQA_B200_LARGEK=1
timeout -s KILL 600 python -m pytest tests/test_gpu_largek.py -m gpu -x -q > gpurun_out/largek_tests.log 2>&1
echo "tests exit $?" | tee -a gpurun_out/largek_tests.log
tail -30 gpurun_out/largek_tests.log
timeout -s KILL 600 compute-sanitizer --tool memcheck python - > gpurun_out/largek_memcheck.log 2>&1 <<'PY'
import numpy as np, proqa_b200 as pq
from tests import data
xb, xq = data.corpus(120_000), data.queries(8)
ix = pq.IndexFlatIP(128); ix.add(xb)
D, I = ix.search(xq, 1500)
print("stats", ix.last_stats)
PY
echo "memcheck exit $?" | tee -a gpurun_out/largek_memcheck.log
tail -15 gpurun_out/largek_memcheck.log
# every path written after round 1's last GPU run, under memcheck and racecheck (includes the merge_lists barrier fix)
export PROQA_B200_STAGED_KMEANS=1
timeout -s KILL 600 python -m pytest tests/test_gpu_sharded_kmeans.py -m gpu -x -q > gpurun_out/staged_kmeans_tests.log 2>&1
echo "staged k-means tests exit $?"; tail -5 gpurun_out/staged_kmeans_tests.log
for tool in memcheck racecheck synccheck; do
  PROQA_B200_SANITIZE_NEW=1 timeout -s KILL 900 compute-sanitizer --tool $tool --error-exitcode 3 python tools/sanitize_small.py > gpurun_out/new_paths_$tool.log 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/new_paths_$tool.log
done
