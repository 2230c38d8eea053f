#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:pq_mma_filter -s 9 -c 1 -o gpurun_out/prof_final3_mma_last python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/final_ncu3.log 2>&1
tail -2 gpurun_out/final_ncu3.log | cut -c1-200
