#!/bin/bash
# Round 2: k = 1 finalize with all loads in flight — k-means parity tests, C4 lines (L2 / IP), the k-means training benchmark.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_kmeans.py tests/test_gpu_sharded_kmeans.py tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -x -q > $O/m_pytest.log 2>&1
echo "gpu tests exit $?"; tail -4 $O/m_pytest.log
for m in l2 ip; do
timeout -s KILL 300 python bench.py --workload c4 --metric $m --steps 5 --warmup 2 --no-cpu-baseline > $O/m_c4_$m.json 2> $O/m_c4_$m.err
python - $O/m_c4_$m.json <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms", round(d["ms_per_step"],2), "points/s", round(d["value"]), "frac", round(d["roofline"]["frac"],3), "filter ms", round(d["roofline"]["kernel_ms_per_step"],2), d["parity"]["ok"])
except Exception as e:
    print("parse failed", e); print(open(sys.argv[1].replace(".json",".err")).read()[-2000:])
PY
done
timeout -s KILL 300 python tools/kmeans_bench.py > $O/m_kmeans.log 2>&1; grep "^\[" $O/m_kmeans.log
