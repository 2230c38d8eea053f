#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests -m gpu -q -x --timeout 150 2>&1 | tail -6 > gpurun_out/final_pytest.log
cat gpurun_out/final_pytest.log
grep -q "passed" gpurun_out/final_pytest.log && ! grep -q "failed\|error" gpurun_out/final_pytest.log || { echo "TESTS FAILED - abort"; exit 1; }
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout -s KILL 300 python bench.py > gpurun_out/final_default.json 2> gpurun_out/final_default.err || { echo "default bench failed"; tail -5 gpurun_out/final_default.err; }
timeout -s KILL 200 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/final_c3.json 2> gpurun_out/final_c3.err
timeout -s KILL 200 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/final_c4.json 2> gpurun_out/final_c4.err
timeout -s KILL 200 python bench.py --workload c5 --rows 12500000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/final_c5shard.json 2> gpurun_out/final_c5shard.err
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:pq_mma_filter -s 5 -c 1 -o gpurun_out/prof_final2_mma_last python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/final_ncu1.log 2>&1
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pq_ -c 200 --csv --log-file gpurun_out/final_launches_c2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/final_ncu2.log 2>&1
for f in final_default final_c3 final_c4 final_c5shard; do python -c "
import json,sys
d=json.load(open('gpurun_out/$f.json'))
print('$f', 'ms',round(d['ms_per_step'],3),'qps',round(d['value'] or 0),'roof',round(d['roofline']['achieved'],1),round(d['roofline']['frac'],3),'parity',d['parity']['ok'],d['parity']['fp32_rerun_queries_per_step'],'kern_ms',round(d['roofline']['kernel_ms_per_step'],3),'e2e',round(d['e2e']['value']),d['clocks'])"; done
