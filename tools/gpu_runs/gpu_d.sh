#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PROQA_B200_MMA=ss
timeout -s KILL 240 python -m pytest tests -m gpu -q -x --timeout 100 -k "bf16 or tiers or golden or eval_fixture or kmeans" 2>&1 | tail -5 > gpurun_out/d_pytest_ss.log
cat gpurun_out/d_pytest_ss.log
grep -q "passed" gpurun_out/d_pytest_ss.log && ! grep -q "failed\|error" gpurun_out/d_pytest_ss.log || { echo "TESTS FAILED - abort"; exit 1; }
timeout -s KILL 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/d_c2_ss.json 2> gpurun_out/d_c2_ss.err || { echo "bench ss failed"; tail -3 gpurun_out/d_c2_ss.err; exit 1; }
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:pq_mma_filter -s 5 -c 1 -o gpurun_out/prof_mma_ss_last python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/d_ncu_ss.log 2>&1
unset PROQA_B200_MMA
timeout -s KILL 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/d_c2_ts.json 2> gpurun_out/d_c2_ts.err
for f in d_c2_ss d_c2_ts; do python -c "
import json,sys
d=json.load(open('gpurun_out/$f.json'))
print('$f', 'ms',round(d['ms_per_step'],3),'qps',d['value'],'roof',round(d['roofline']['achieved'],1),round(d['roofline']['frac'],3),'parity',d['parity']['ok'],d['parity']['fp32_rerun_queries_per_step'],'kern_ms',d['roofline']['kernel_ms_per_step'],d['clocks'])"; tail -2 gpurun_out/$f.err; done
