#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 100 -k "bf16 or tiers or eval or kmeans" 2>&1 | tail -3 > gpurun_out/m_pytest.log
cat gpurun_out/m_pytest.log
grep -q "passed" gpurun_out/m_pytest.log && ! grep -q "failed\|error" gpurun_out/m_pytest.log || { echo "TESTS FAILED - abort"; exit 1; }
for i in 1 2; do timeout -s KILL 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/m_c2_$i.json 2> gpurun_out/m_c2.err || { echo "bench failed"; tail -3 gpurun_out/m_c2.err; exit 1; }; done
for f in m_c2_1 m_c2_2; do python -c "
import json,sys
d=json.load(open('gpurun_out/$f.json'))
print('$f', 'ms',round(d['ms_per_step'],3),'qps',d['value'],'roof',round(d['roofline']['achieved'],1),round(d['roofline']['frac'],3),'parity',d['parity']['ok'],d['parity']['fp32_rerun_queries_per_step'],'kern_ms',d['roofline']['kernel_ms_per_step'],d['clocks'])"; done
