#!/bin/bash
# ncu full-set capture of the last (tight-threshold) epoch of a C2 search with the one-vote epilogue
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:pq_mma_filter -s 5 -c 1 -o gpurun_out/x_prof_last -f python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline --no-sweep > gpurun_out/x_ncu.log 2>&1
echo "ncu rc=$?"
