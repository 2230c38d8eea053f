#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests -m gpu -q -x --timeout 150 2>&1 | tail -15 > gpurun_out/r_pytest.log
cat gpurun_out/r_pytest.log
grep -q "passed" gpurun_out/r_pytest.log && ! grep -q "failed\|error" gpurun_out/r_pytest.log || { echo "TESTS FAILED - abort"; exit 1; }
timeout -s KILL 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r_c2.json 2> gpurun_out/r_c2.err || { echo "bench failed"; tail -3 gpurun_out/r_c2.err; exit 1; }
timeout -s KILL 200 python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r_c4.json 2> gpurun_out/r_c4.err
for f in r_c2 r_c4; do python -c "
import json,sys
d=json.load(open('gpurun_out/$f.json'))
print('$f', 'ms',round(d['ms_per_step'],3),'qps',round(d['value'] or 0),'roof',round(d['roofline']['achieved'],1),round(d['roofline']['frac'],3),'parity',d['parity']['ok'],d['parity']['fp32_rerun_queries_per_step'],'kern_ms',round(d['roofline']['kernel_ms_per_step'],3),d['clocks']['sm_mhz'])"; done
