#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:pq_mma_filter -s 1 -c 3 -o gpurun_out/prof_mma_small_epochs python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/k_ncu.log 2>&1
tail -2 gpurun_out/k_ncu.log | cut -c1-200
