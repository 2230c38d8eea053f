#!/bin/bash
# Final build: the k-means assignment pass at full size (C4: 21M points x 10,000 centroids, k = 1, L2 and IP) and the training bench.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
for m in l2 ip; do
timeout -s KILL 300 python bench.py --workload c4 --metric $m --steps 5 --warmup 2 --no-cpu-baseline > $O/zh_c4_$m.json 2> $O/zh_c4_$m.err
python - $O/zh_c4_$m.json <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms", round(d["ms_per_step"],3), "value", round(d["value"]), d["unit"], "filter ms", round(d["roofline"]["kernel_ms_per_step"],2), "frac", round(d["roofline"]["frac"],3), d["parity"]["ok"], "e2e", round(d["e2e"]["ms_per_step"],2), d.get("clocks"))
except Exception as e:
    print("parse failed", sys.argv[1], e); print(open(sys.argv[1].replace(".json",".err")).read()[-2000:])
PY
done
timeout -s KILL 300 python tools/kmeans_bench.py > $O/zh_kmeans.log 2>&1; tail -6 $O/zh_kmeans.log
