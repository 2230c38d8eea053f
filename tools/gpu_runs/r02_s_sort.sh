#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_kmeans.py tests/test_gpu_sharded_kmeans.py tests/test_gpu_multi.py -m gpu -x -q > $O/s_pytest.log 2>&1
echo "gpu tests exit $?"; tail -4 $O/s_pytest.log
timeout -s KILL 300 python tools/kmeans_bench.py > $O/s_kmeans.log 2>&1; grep "^\[" $O/s_kmeans.log
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:km_ -c 60 --csv --log-file $O/s_launches_km.csv python tools/kmeans_bench.py 2000000 > $O/s_ncu.log 2>&1
python - <<'PY'
import csv, collections
with open('gpurun_out/s_launches_km.csv') as f:
    lines=[l for l in f if not l.startswith('==')]
agg=collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get('Metric Name')!='gpu__time_duration.sum': continue
    name=row['Kernel Name'].split('(')[0]
    val=float(row['Metric Value'].replace(',','')); unit=row['Metric Unit']
    agg.setdefault(name,[]).append(val/1000 if unit in ('ns','nsecond') else val)
for n,us in agg.items(): print(f"{n:40s} n={len(us):3d} mean_us={sum(us)/len(us):9.1f}")
PY
