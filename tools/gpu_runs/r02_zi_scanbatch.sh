#!/bin/bash
# fp32 scan with all query batches of a search in one launch: parity tests, smoke (160 queries on the fp32 tier = 20 batches), the
# k-means training bench (iteration 2 on random data re-runs ~10^5 points through the scan).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kmeans.py tests/test_gpu_largek.py tests/test_gpu_merge.py -m gpu -x -q > $O/zi_pytest.log 2>&1
echo "gpu tests exit $?"; tail -3 $O/zi_pytest.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/zi_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/zi_smoke.log
timeout -s KILL 300 python tools/kmeans_bench.py > $O/zi_kmeans.log 2>&1; tr '\r' '\n' < $O/zi_kmeans.log | grep -v "^$" | tail -16
timeout -s KILL 100 python - <<'PY'
import time, numpy as np, torch, proqa_b200 as pq
rng = np.random.default_rng(0)
xb = rng.standard_normal((10_000, 128), dtype=np.float32)
xq = rng.standard_normal((100_000, 128), dtype=np.float32)
ix = pq.IndexFlatL2(128); ix.add(xb); ix.set_tier("fp32")
for _ in range(2):
    t = time.time(); D, I = ix.search(xq, 1); dt = time.time() - t
print("fp32 scan, 100k queries x 10k rows, k=1 (12.5k batches):", round(dt * 1e3, 1), "ms host wall, stats", ix.last_stats[:8])
ix.set_tier("auto"); D2, I2 = ix.search(xq, 1)
print("same bits as the tensor tier:", bool((I == I2).all() and (D.view(np.uint32) == D2.view(np.uint32)).all()))
PY
