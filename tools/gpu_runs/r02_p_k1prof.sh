#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:pq_k1_finalize -s 2 -c 1 -o $O/p_prof_k1 -f python bench.py --workload c4 --nq 2097152 --steps 1 --warmup 1 --no-cpu-baseline > $O/p_ncu.log 2>&1
echo "ncu rc=$?"
