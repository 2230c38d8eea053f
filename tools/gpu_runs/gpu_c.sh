#!/bin/bash
# fail fast: a hung kernel must not eat the GPU budget
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 240 python -m pytest tests -m gpu -q -x --timeout 100 2>&1 | tail -15 > gpurun_out/c_pytest.log
cat gpurun_out/c_pytest.log
grep -q "passed" gpurun_out/c_pytest.log && ! grep -q "failed\|error" gpurun_out/c_pytest.log || { echo "TESTS FAILED - abort"; exit 1; }
timeout -s KILL 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench_c2.json 2> gpurun_out/c_bench_c2.err || { echo "bench c2 failed"; tail -3 gpurun_out/c_bench_c2.err; exit 1; }
timeout -s KILL 200 python bench.py --workload s0 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c_s0_auto.json 2> gpurun_out/c_s0_auto.err
timeout -s KILL 200 python bench.py --workload c1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c_c1.json 2> gpurun_out/c_c1.err
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:pq_mma_filter -s 5 -c 1 -o gpurun_out/prof_mma_ts_last python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c_ncu_mma.log 2>&1
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pq_ -c 200 --csv --log-file gpurun_out/c_launches_c2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c_ncu_launch.log 2>&1
for f in c_bench_c2 c_s0_auto c_c1; do python -c "
import json,sys
d=json.load(open('gpurun_out/$f.json'))
print('$f', 'ms',round(d['ms_per_step'],3),'qps',d['value'],'roof',round(d['roofline']['achieved'],1),round(d['roofline']['frac'],3),'parity',d['parity']['ok'],d['parity']['fp32_rerun_queries_per_step'],'kern_ms',d['roofline']['kernel_ms_per_step'])"; tail -2 gpurun_out/$f.err; done
