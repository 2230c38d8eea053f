#!/bin/bash
# Round 2, GPU call 2: warp-per-query select/rescore, deferred repair, no per-epoch memsets — parity, then the C2 / S0 / C1 lines.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > $O/b_pytest.log 2>&1
echo "gpu tests exit $?"; tail -6 $O/b_pytest.log
for wl in c2 c1; do
  timeout -s KILL 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline > $O/b_$wl.json 2> $O/b_$wl.err
  echo "$wl rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$O/b_$wl.json").read().strip().splitlines()[-1])
    print("$wl", d["ms_per_step"], d["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_step"], d["gpu_launches"], d["parity"], d["e2e"]["ms_per_step"])
except Exception as e:
    print("parse failed", e); print(open("$O/b_$wl.err").read()[-2000:])
PY
done
for nq in 1 16 64 256; do
  timeout -s KILL 200 python bench.py --workload s0 --nq $nq --steps 20 --warmup 3 --no-cpu-baseline > $O/b_s0_$nq.json 2> $O/b_s0_$nq.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/b_s0_$nq.json").read().strip().splitlines()[-1])
    print("s0 nq=$nq", d["ms_per_step"], d["roofline"]["bound"], d["roofline"]["frac"], d["parity"]["ok"])
except Exception as e:
    print("parse failed", e); print(open("$O/b_s0_$nq.err").read()[-1500:])
PY
done
# launch list of one C2 step (shares) 
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pq_ -c 120 --csv --log-file $O/b_launches_c2.csv python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline > $O/b_ncu_launch.log 2>&1
echo "ncu rc=$?"
