#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pq_ -s 46 -c 70 --csv --log-file gpurun_out/g_launches_c4.csv python bench.py --workload c4 --metric l2 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/g_ncu_launch.log 2>&1
tail -1 gpurun_out/g_ncu_launch.log | cut -c1-300
