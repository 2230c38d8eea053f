#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests -m gpu -q -x --timeout 120 2>&1 | tail -12 > gpurun_out/q_pytest.log
cat gpurun_out/q_pytest.log
timeout -s KILL 200 python bench.py --workload s0 --nq 1 --k 5000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/q_nq1_k5000.json 2> gpurun_out/q_nq1_k5000.err || tail -3 gpurun_out/q_nq1_k5000.err
python -c "
import json
d=json.load(open('gpurun_out/q_nq1_k5000.json'))
print('nq1 k5000', 'ms',round(d['ms_per_step'],3),'roof',round(d['roofline']['achieved'],1),d['roofline']['unit'],round(d['roofline']['frac'],3),'parity',d['parity']['ok'])"
