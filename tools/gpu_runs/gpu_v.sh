#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 100 -k "bf16 or tiers or eval or kmeans or document" 2>&1 | tail -3 > gpurun_out/v_pytest.log
cat gpurun_out/v_pytest.log
grep -q "passed" gpurun_out/v_pytest.log && ! grep -q "failed\|error" gpurun_out/v_pytest.log || { echo "TESTS FAILED - abort"; exit 1; }
timeout -s KILL 200 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/v_c3.json 2> gpurun_out/v_c3.err
timeout -s KILL 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/v_c2.json 2> gpurun_out/v_c2.err
for f in v_c3 v_c2; do python -c "
import json,sys
d=json.load(open('gpurun_out/$f.json'))
print('$f', 'ms',round(d['ms_per_step'],3),'qps',round(d['value'] or 0),'roof',round(d['roofline']['achieved'],1),round(d['roofline']['frac'],3),'parity',d['parity']['ok'],d['parity']['fp32_rerun_queries_per_step'],'kern_ms',round(d['roofline']['kernel_ms_per_step'],3),d['clocks']['sm_mhz'])"; done
