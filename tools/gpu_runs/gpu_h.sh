#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
G=${1:-2}
run() { name=$1; shift; timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err || { echo "$name failed"; tail -8 gpurun_out/$name.err; }; }
run p_c2_rows_g$G --steps 20 --warmup 3 --no-cpu-baseline
run p_c2_q_g$G --steps 20 --warmup 3 --no-cpu-baseline --row-shards 1
run p_c3_rows_g$G --workload c3 --steps 3 --warmup 3 --no-cpu-baseline
run p_c3_q_g$G --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --row-shards 1
for f in p_c2_rows_g$G p_c2_q_g$G p_c3_rows_g$G p_c3_q_g$G; do python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/$f.json') if l.startswith('{')][-1])
print('$f', d['config']['parallelism'], 'ms',round(d['ms_per_step'],3),'qps',round(d['value'] or 0),'roof',round(d['roofline']['achieved'],1),round(d['roofline']['frac'],3),'parity',d['parity']['ok'],'kern_ms',round(d['roofline']['kernel_ms_per_step'],2),'e2e',round(d['e2e']['value']),'phases',d['multi_gpu_phases'])"; done
