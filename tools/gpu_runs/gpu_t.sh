#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 150 -k "bf16 or tiers or eval or document or certificate" 2>&1 | tail -4 > gpurun_out/t_pytest.log
cat gpurun_out/t_pytest.log
grep -q "passed" gpurun_out/t_pytest.log && ! grep -q "failed\|error" gpurun_out/t_pytest.log || { echo "TESTS FAILED - abort"; exit 1; }
for nq in 16 64 256 1024; do
  timeout -s KILL 200 python bench.py --workload s0 --nq $nq --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/t_s0_nq$nq.json 2> gpurun_out/t_s0_nq$nq.err || { echo "s0 $nq failed"; tail -3 gpurun_out/t_s0_nq$nq.err; }
done
for f in t_s0_nq16 t_s0_nq64 t_s0_nq256 t_s0_nq1024; do python -c "
import json,sys
d=json.load(open('gpurun_out/$f.json'))
print('$f', 'ms',round(d['ms_per_step'],3),'qps',round(d['value'] or 0,1),'corpusGB/s',round(d['corpus_gbs'],1),round(d['roofline']['achieved'],1),d['roofline']['unit'],round(d['roofline']['frac'],3),'parity',d['parity']['ok'],d['parity']['fp32_rerun_queries_per_step'],'launches/step',d['gpu_launches']/20)"; done
