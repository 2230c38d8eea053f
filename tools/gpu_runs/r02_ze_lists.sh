#!/bin/bash
# Launch lists (ncu, durations only) of the C5 shard (8192 q x 12.5M, k=1000) and the trec shape (256 q x 8.8M, k=10000), plus the
# new multi-wave parity test.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "waves or online_sampler" > $O/ze_pytest.log 2>&1
echo "tests exit $?"; tail -3 $O/ze_pytest.log
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pq_ -c 120 --csv --log-file $O/ze_launches_c5.csv python bench.py --workload c5 --rows 12500000 --steps 1 --warmup 1 --no-cpu-baseline > $O/ze_ncu_c5.log 2>&1
echo "ncu c5 rc=$?"
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pq_ -c 120 --csv --log-file $O/ze_launches_trec.csv python bench.py --workload trec --steps 1 --warmup 1 --no-cpu-baseline > $O/ze_ncu_trec.log 2>&1
echo "ncu trec rc=$?"
python - <<'PY'
import csv, re
for name in ("c5", "trec"):
    rows = [r for r in csv.reader(open(f"gpurun_out/ze_launches_{name}.csv", errors="replace")) if len(r) > 10 and r[0].isdigit()]
    L = [(re.sub(r"^void ", "", re.sub(r"\(.*", "", r[4])), r[8], float(r[-1].replace(",", "")) / 1e3) for r in rows]
    starts = [i for i, l in enumerate(L) if l[0].startswith("pq::pq_mma_init_state_kernel") or l[0].startswith("pq_mma_init_state_kernel")]
    print("==", name, "launches", len(L), "searches", len(starts))
    if len(starts) >= 2:
        s, e = starts[-2], starts[-1]
        for l in L[s:e]: print("  ", l[0][:60], l[1], round(l[2], 1))
        print("   total us", round(sum(l[2] for l in L[s:e]), 1))
PY
