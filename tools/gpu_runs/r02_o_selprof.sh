#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:pq_epoch_select_warp -s 9 -c 1 -o $O/o_prof_select -f python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline --no-sweep > $O/o_ncu_sel.log 2>&1
echo "ncu rc=$?"
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:pq_rescore_warp -s 1 -c 1 -o $O/o_prof_rescore -f python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline --no-sweep > $O/o_ncu_resc.log 2>&1
echo "ncu rc=$?"
