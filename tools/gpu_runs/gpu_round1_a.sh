#!/bin/bash
# first measured pass: bench lines (c2 auto, s0 fp32 / auto), launch list and one full ncu capture of the filter kernel
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
timeout 300 python bench.py --workload s0 --tier fp32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s0_fp32.json 2> gpurun_out/bench_s0_fp32.err
timeout 300 python bench.py --workload s0 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s0_auto.json 2> gpurun_out/bench_s0_auto.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pq_ -c 200 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pq_mma_filter -s 6 -c 2 -o gpurun_out/prof_mma_c2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/bench_c2.json gpurun_out/bench_s0_fp32.json gpurun_out/bench_s0_auto.json
tail -3 gpurun_out/*.err
