#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -x -q > $O/q_pytest.log 2>&1
echo "gpu tests exit $?"; tail -4 $O/q_pytest.log
for m in l2 ip; do
timeout -s KILL 300 python bench.py --workload c4 --metric $m --steps 5 --warmup 2 --no-cpu-baseline > $O/q_c4_$m.json 2> $O/q_c4_$m.err
python - $O/q_c4_$m.json <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms", round(d["ms_per_step"],2), "points/s", round(d["value"]), "frac", round(d["roofline"]["frac"],3), "filter ms", round(d["roofline"]["kernel_ms_per_step"],2), d["parity"]["ok"])
except Exception as e:
    print("parse failed", e); print(open(sys.argv[1].replace(".json",".err")).read()[-2000:])
PY
done
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sweep > $O/q_c2.json 2> $O/q_c2.err
python - $O/q_c2.json <<'PY'
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("c2 ms", round(d["ms_per_step"],3), "frac", round(d["roofline"]["frac"],4), d["roofline"]["other_kernels_ms_per_step"], d["parity"]["ok"], d["clocks"])
PY
timeout -s KILL 300 python tools/kmeans_bench.py > $O/q_kmeans.log 2>&1; grep "^\[" $O/q_kmeans.log
