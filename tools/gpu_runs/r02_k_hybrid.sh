#!/bin/bash
# Round 2: kernel variant per epoch (4 epilogue sets while the threshold is loose, 2 once it is tight) — parity, C2 / C3 lines.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -x -q > $O/k_pytest.log 2>&1
echo "gpu tests exit $?"; tail -5 $O/k_pytest.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms", round(d["ms_per_step"],3), "q/s", d["value"] and round(d["value"]), "frac", round(d["roofline"]["frac"],4), "kern ms", round(d["roofline"]["kernel_ms_per_step"],3), d["roofline"].get("other_kernels_ms_per_step"),
          "launches", d["gpu_launches"], "parity", d["parity"]["ok"], "e2e ms", round(d["e2e"]["ms_per_step"],3), d.get("clocks"))
except Exception as e:
    print("parse failed", sys.argv[1], e); print(open(sys.argv[1].replace(".json",".err")).read()[-3000:])
PY
}
for i in 1 2; do
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sweep > $O/k_c2_$i.json 2> $O/k_c2_$i.err; show $O/k_c2_$i.json
done
timeout -s KILL 400 python bench.py --workload c3 --steps 5 --warmup 2 --no-cpu-baseline > $O/k_c3.json 2> $O/k_c3.err; show $O/k_c3.json
timeout -s KILL 300 python bench.py --workload c1 --steps 20 --warmup 5 --no-cpu-baseline > $O/k_c1.json 2> $O/k_c1.err; show $O/k_c1.json
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pq_ -c 60 --csv --log-file $O/k_launches_c2.csv python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline --no-sweep > $O/k_ncu_launch.log 2>&1
echo "ncu launch rc=$?"
