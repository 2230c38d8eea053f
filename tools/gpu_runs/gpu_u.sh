#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PROQA_B200_N128=1
timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 100 -k "bf16 or tiers or eval or kmeans" 2>&1 | tail -3 > gpurun_out/u_pytest.log
cat gpurun_out/u_pytest.log
grep -q "passed" gpurun_out/u_pytest.log && ! grep -q "failed\|error" gpurun_out/u_pytest.log || { echo "TESTS FAILED - abort"; exit 1; }
for i in 1 2 3; do
  PROQA_B200_N128=1 timeout -s KILL 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/u_c2_n128_$i.json 2> gpurun_out/u.err
  PROQA_B200_N128=0 timeout -s KILL 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/u_c2_n64_$i.json 2> gpurun_out/u.err
done
for f in u_c2_n128_1 u_c2_n64_1 u_c2_n128_2 u_c2_n64_2 u_c2_n128_3 u_c2_n64_3; do python -c "
import json,sys
d=json.load(open('gpurun_out/$f.json'))
print('$f', 'ms',round(d['ms_per_step'],3),'qps',round(d['value'] or 0),'roof',round(d['roofline']['achieved'],1),round(d['roofline']['frac'],3),'parity',d['parity']['ok'],'kern_ms',round(d['roofline']['kernel_ms_per_step'],3),d['clocks']['sm_mhz'])"; done
