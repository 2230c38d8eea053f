#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pq_ -s 46 -c 80 --csv --log-file gpurun_out/g_launches_c4.csv python bench.py --workload c4 --metric l2 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/g_ncu_launch.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:pq_mma_filter -s 25 -c 1 -o gpurun_out/prof_mma_c4 python bench.py --workload c4 --metric l2 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/g_ncu_full.log 2>&1
tail -2 gpurun_out/g_ncu_full.log | cut -c1-200
