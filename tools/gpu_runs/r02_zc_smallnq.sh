#!/bin/bash
# Small batches: nq = 1 / 4 through the tensor tier (bf16 copy, 256 B per row) against the fp32 scan (512 B per row), k = 80, 1000
# and 5000 (online_sampler.py:113); launch lists of S0 nq = 16 and C1 (where the fixed costs of the epochs sit).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
show() { python - "$1" <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms", round(d["ms_per_step"],3), "roofline", d["roofline"]["bound"], round(d["roofline"]["frac"],4), "kern ms", round(d["roofline"]["kernel_ms_per_step"],3),
          "parity", d["parity"]["ok"], "e2e ms", round(d["e2e"]["ms_per_step"],3), "launches", d["gpu_launches"])
except Exception as e:
    print("parse failed", sys.argv[1], e); print(open(sys.argv[1].replace(".json",".err")).read()[-2000:])
PY
}
for k in 80 1000 5000; do
for nq in 1 4; do
for mq in 5 1; do
PROQA_B200_MMA_MIN_QUERIES=$mq timeout -s KILL 200 python bench.py --workload s0 --nq $nq --k $k --steps 20 --warmup 3 --no-cpu-baseline > $O/zc_s0_nq${nq}_k${k}_mq$mq.json 2> $O/zc_s0_nq${nq}_k${k}_mq$mq.err; show $O/zc_s0_nq${nq}_k${k}_mq$mq.json
done
done
done
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pq_ -c 80 --csv --log-file $O/zc_launches_s0_nq16.csv python bench.py --workload s0 --steps 1 --warmup 1 --no-cpu-baseline > $O/zc_ncu_s0.log 2>&1
echo "ncu s0 rc=$?"
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pq_ -c 80 --csv --log-file $O/zc_launches_c1.csv python bench.py --workload c1 --steps 1 --warmup 1 --no-cpu-baseline > $O/zc_ncu_c1.log 2>&1
echo "ncu c1 rc=$?"
python tools/launch_shares.py $O/zc_launches_s0_nq16.csv
python tools/launch_shares.py $O/zc_launches_c1.csv
