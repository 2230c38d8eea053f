#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
for sets in 2 4; do for m in l2 ip; do
PROQA_B200_K1_SETS=$sets timeout -s KILL 300 python bench.py --workload c4 --metric $m --steps 5 --warmup 2 --no-cpu-baseline > $O/r_c4_${m}_$sets.json 2> $O/r_c4_${m}_$sets.err
python - $O/r_c4_${m}_$sets.json <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms", round(d["ms_per_step"],2), "frac", round(d["roofline"]["frac"],3), "filter ms", round(d["roofline"]["kernel_ms_per_step"],2), d["parity"]["ok"])
except Exception as e:
    print("parse failed", e); print(open(sys.argv[1].replace(".json",".err")).read()[-2000:])
PY
done; done
