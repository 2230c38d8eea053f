#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -15 > gpurun_out/b_pytest.log
for nq in 1 4 8; do
  timeout 300 python bench.py --workload s0 --nq $nq --tier fp32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b_s0_fp32_nq$nq.json 2> gpurun_out/b_s0_fp32_nq$nq.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pq_ffma_scan -s 1 -c 1 -o gpurun_out/prof_ffma_nq8 python bench.py --workload s0 --nq 8 --tier fp32 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu_ffma.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pq_mma_filter -s 5 -c 1 -o gpurun_out/prof_mma_c2_last python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu_mma.log 2>&1
cat gpurun_out/b_pytest.log
for nq in 1 4 8; do python -c "
import json,sys
d=json.load(open('gpurun_out/b_s0_fp32_nq$nq.json'))
print('nq=$nq', 'ms',d['ms_per_step'],'roof',d['roofline']['achieved'],d['roofline']['frac'],'parity',d['parity']['ok'])"; done
