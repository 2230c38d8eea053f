#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests -m gpu -q -x --timeout 120 2>&1 | tail -5 > gpurun_out/j_pytest.log
cat gpurun_out/j_pytest.log
grep -q "passed" gpurun_out/j_pytest.log && ! grep -q "failed\|error" gpurun_out/j_pytest.log || { echo "TESTS FAILED - abort"; exit 1; }
timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/j_c2.json 2> gpurun_out/j_c2.err || { echo "bench failed"; tail -3 gpurun_out/j_c2.err; exit 1; }
timeout -s KILL 200 python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/j_c3.json 2> gpurun_out/j_c3.err
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pq_ -c 120 --csv --log-file gpurun_out/j_launches_c2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/j_ncu_launch.log 2>&1
for f in j_c2 j_c3; do python -c "
import json,sys
d=json.load(open('gpurun_out/$f.json'))
print('$f', 'ms',round(d['ms_per_step'],3),'qps',d['value'],'roof',round(d['roofline']['achieved'],1),round(d['roofline']['frac'],3),'parity',d['parity']['ok'],d['parity']['fp32_rerun_queries_per_step'],'kern_ms',d['roofline']['kernel_ms_per_step'],d['clocks'])"; done
