#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
run() { name=$1; shift; timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err || { echo "$name failed"; tail -8 gpurun_out/$name.err; }; }
run h_c2_g2 --steps 10 --warmup 3 --no-cpu-baseline
for f in h_c2_g2; do python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/$f.json') if l.startswith('{')][-1])
print('$f', 'ms',round(d['ms_per_step'],3),'qps',d['value'],'roof',round(d['roofline']['achieved'],1),round(d['roofline']['frac'],3),'parity',d['parity']['ok'],'kern_ms',d['roofline']['kernel_ms_per_step'],'e2e',d['e2e'],'phases',d['multi_gpu_phases'])"; done
