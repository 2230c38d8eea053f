#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:pq_largek_finalize -s 1 -c 1 -o gpurun_out/u_prof_fin -f python bench.py --workload trec --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/u_ncu.log 2>&1
echo "ncu rc=$?"
