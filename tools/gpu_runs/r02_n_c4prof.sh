#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/n_launches_c4.csv python bench.py --workload c4 --nq 2097152 --steps 1 --warmup 1 --no-cpu-baseline > $O/n_ncu.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import csv
with open('gpurun_out/n_launches_c4.csv') as f:
    lines=[l for l in f if not l.startswith('==')]
seq=[]
for row in csv.DictReader(lines):
    if row.get('Metric Name')!='gpu__time_duration.sum': continue
    name=row['Kernel Name'].split('(')[0]
    val=float(row['Metric Value'].replace(',','')); unit=row['Metric Unit']
    seq.append((name, val/1000 if unit in ('ns','nsecond') else val))
for n,u in seq[-40:]: print(f"{u:10.1f} us  {n}")
PY
