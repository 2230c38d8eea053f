#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 300 python bench.py --workload c5 --rows 12500000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/l_c5shard.json 2> gpurun_out/l_c5shard.err || { echo "c5 failed"; tail -5 gpurun_out/l_c5shard.err; }
for nq in 1 4 8 16 64 256 1024; do
  timeout -s KILL 200 python bench.py --workload s0 --nq $nq --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/l_s0_nq$nq.json 2> gpurun_out/l_s0_nq$nq.err || { echo "s0 $nq failed"; tail -3 gpurun_out/l_s0_nq$nq.err; }
done
for f in l_c5shard l_s0_nq1 l_s0_nq4 l_s0_nq8 l_s0_nq16 l_s0_nq64 l_s0_nq256 l_s0_nq1024; do python -c "
import json,sys
d=json.load(open('gpurun_out/$f.json'))
print('$f', 'ms',round(d['ms_per_step'],3),'qps',round(d['value'] or 0,1),'corpusGB/s',round(d['corpus_gbs'],1),d['roofline']['kernel'],round(d['roofline']['achieved'],1),d['roofline']['unit'],round(d['roofline']['frac'],3),'parity',d['parity']['ok'],d['parity']['fp32_rerun_queries_per_step'],'e2e_ms',round(d['e2e']['ms_per_step'],3))"; done
