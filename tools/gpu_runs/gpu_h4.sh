#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
G=4
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --steps 20 --warmup 3 > gpurun_out/v_c2_auto_g4.json 2> gpurun_out/v_c2_auto_g4.err || tail -5 gpurun_out/v_c2_auto_g4.err
python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/v_c2_auto_g4.json') if l.startswith('{')][-1])
print('g4', d['config']['parallelism'], 'ms',round(d['ms_per_step'],3),'qps',round(d['value'] or 0),'roof',round(d['roofline']['frac'],3),'parity',d['parity']['ok'],'e2e',round(d['e2e']['value']))"
