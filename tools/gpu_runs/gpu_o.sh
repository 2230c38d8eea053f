#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout -s KILL 400 compute-sanitizer --tool $tool --error-exitcode 3 python tools/sanitize_small.py > gpurun_out/o_san_$tool.log 2>&1
  echo "$tool rc=$?"; tail -4 gpurun_out/o_san_$tool.log
done
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:pq_mma_filter -s 5 -c 1 -o gpurun_out/prof_final_mma_last python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/o_ncu1.log 2>&1
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pq_ -c 150 --csv --log-file gpurun_out/o_launches_c2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/o_ncu2.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:pq_ffma_scan -s 1 -c 1 -o gpurun_out/prof_final_ffma_nq4 python bench.py --workload s0 --nq 4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/o_ncu3.log 2>&1
ls -la gpurun_out/prof_final* gpurun_out/o_launches_c2.csv
