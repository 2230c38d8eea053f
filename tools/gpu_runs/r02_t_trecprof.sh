#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pq_ -c 80 --csv --log-file $O/t_launches_trec.csv python bench.py --workload trec --steps 1 --warmup 1 --no-cpu-baseline > $O/t_ncu.log 2>&1
python - <<'PY'
import csv
with open('gpurun_out/t_launches_trec.csv') as f:
    lines=[l for l in f if not l.startswith('==')]
seq=[]
for row in csv.DictReader(lines):
    if row.get('Metric Name')!='gpu__time_duration.sum': continue
    name=row['Kernel Name'].split('(')[0]
    val=float(row['Metric Value'].replace(',','')); unit=row['Metric Unit']
    seq.append((name, val/1000 if unit in ('ns','nsecond') else val))
for n,u in seq[-24:]: print(f"{u:10.1f} us  {n}")
PY
