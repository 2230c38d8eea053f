#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 200 python bench.py --workload c4 --metric ip --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/w_c4_ip.json 2> gpurun_out/w_c4_ip.err
for nq in 1 4 16 64 256 1024; do
  timeout -s KILL 200 python bench.py --workload s0 --nq $nq --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/w_s0_nq$nq.json 2> gpurun_out/w_s0_nq$nq.err || { echo "s0 $nq failed"; tail -3 gpurun_out/w_s0_nq$nq.err; }
done
for f in w_c4_ip w_s0_nq1 w_s0_nq4 w_s0_nq16 w_s0_nq64 w_s0_nq256 w_s0_nq1024; do python -c "
import json,sys
d=json.load(open('gpurun_out/$f.json'))
print('$f', 'ms',round(d['ms_per_step'],3),'qps',round(d['value'] or 0,1),'corpusGB/s',round(d['corpus_gbs'],1),d['roofline']['kernel'],round(d['roofline']['achieved'],1),d['roofline']['unit'],round(d['roofline']['frac'],3),'parity',d['parity']['ok'],'e2e',round(d['e2e']['value'],1))"; done
