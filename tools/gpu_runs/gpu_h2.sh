#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
G=${1:-2}
run() { name=$1; shift; timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err || { echo "$name failed"; tail -8 gpurun_out/$name.err; }; }
run v_c2_auto_g$G --steps 20 --warmup 3 --no-cpu-baseline
run v_c2_rows_g$G --steps 20 --warmup 3 --no-cpu-baseline --row-shards rows
timeout -s KILL 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --impl reference --steps 1 --warmup 0 2>/dev/null | cut -c1-200
for f in v_c2_auto_g$G v_c2_rows_g$G; do python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/$f.json') if l.startswith('{')][-1])
print('$f', d['config']['parallelism'], 'ms',round(d['ms_per_step'],3),'qps',round(d['value'] or 0),'roof',round(d['roofline']['frac'],3),'parity',d['parity']['ok'],'e2e',round(d['e2e']['value']),'phases',d['multi_gpu_phases'])"; done
