#!/bin/bash
# Which epochs run the four-set (loose) filter variant: threshold on the expected share of 32x32 chunks with a survivor, A/B on C2.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
show() { python - "$1" <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms", round(d["ms_per_step"],3), "frac", round(d["roofline"]["frac"],4), "kern ms", round(d["roofline"]["kernel_ms_per_step"],3), "parity", d["parity"]["ok"], d.get("clocks"))
except Exception as e:
    print("parse failed", sys.argv[1], e); print(open(sys.argv[1].replace(".json",".err")).read()[-2000:])
PY
}
for la in 0.10 0.35 1.5 0.10 0.35 1.5; do
PROQA_B200_LOOSE_ABOVE=$la timeout -s KILL 300 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline --no-sweep > $O/zg_c2_la$la.json 2> $O/zg_c2_la$la.err; show $O/zg_c2_la$la.json
done
