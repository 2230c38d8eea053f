#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 120 python -m pytest tests -m gpu -q -x --timeout 100 -k "kmeans or l2" 2>&1 | tail -3
timeout -s KILL 150 python bench.py --workload c4 --metric l2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/e_c4_l2.json 2> gpurun_out/e_c4_l2.err || { echo "c4 l2 failed"; tail -5 gpurun_out/e_c4_l2.err; }
timeout -s KILL 150 python bench.py --workload c4 --metric ip --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/e_c4_ip.json 2> gpurun_out/e_c4_ip.err || { echo "c4 ip failed"; tail -5 gpurun_out/e_c4_ip.err; }
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pq_ -s 400 -c 400 --csv --log-file gpurun_out/f_launches_c4.csv python bench.py --workload c4 --metric l2 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/f_ncu_launch.log 2>&1
for f in e_c4_l2 e_c4_ip; do python -c "
import json,sys
d=json.load(open('gpurun_out/$f.json'))
print('$f', 'ms',round(d['ms_per_step'],3),'qps',d['value'],'roof',round(d['roofline']['achieved'],1),round(d['roofline']['frac'],3),'parity',d['parity'],'kern_ms',d['roofline']['kernel_ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])"; tail -2 gpurun_out/$f.err; done
