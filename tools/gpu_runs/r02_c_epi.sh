#!/bin/bash
# Round 2, GPU call 3: 16 epilogue warps, select dispatch, one-process multi-GPU index (shards on one device here), sweep + trec lines.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > $O/c_pytest.log 2>&1
echo "gpu tests exit $?"; tail -8 $O/c_pytest.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms", round(d["ms_per_step"],3), "q/s", d["value"] and round(d["value"]), "frac", round(d["roofline"]["frac"],4), "kern ms", round(d["roofline"]["kernel_ms_per_step"],3),
          "launches", d["gpu_launches"], "parity", d["parity"]["ok"], d["parity"].get("max_near_tie_gap_rel"), "e2e ms", round(d["e2e"]["ms_per_step"],3), d["e2e"].get("pinned_buffers_ms_per_step"), d.get("clocks"))
    for k_, v in (d.get("sweep") or {}).items(): print("   ", k_, {a: (round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a != "step_frac_of_hbm_note"})
except Exception as e:
    print("parse failed", sys.argv[1], e); print(open(sys.argv[1].replace(".json",".err")).read()[-3000:])
PY
}
timeout -s KILL 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/c_c2.json 2> $O/c_c2.err; show $O/c_c2.json
PROQA_B200_LIB=$PWD/proqa_b200/libproqa_b200_epi8.so timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-sweep > $O/c_c2_epi8.json 2> $O/c_c2_epi8.err; show $O/c_c2_epi8.json
timeout -s KILL 300 python bench.py --workload trec --steps 10 --warmup 3 > $O/c_trec.json 2> $O/c_trec.err; show $O/c_trec.json
timeout -s KILL 300 python bench.py --workload c4 --steps 5 --warmup 2 --no-cpu-baseline > $O/c_c4.json 2> $O/c_c4.err; show $O/c_c4.json
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pq_ -c 80 --csv --log-file $O/c_launches_c2.csv python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline --no-sweep > $O/c_ncu_launch.log 2>&1
echo "ncu launch rc=$?"
# full captures: 4th filter launch of a search (rows 64k..512k) and the last one
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:pq_mma_filter -s 3 -c 1 -o $O/c_prof_epoch3 -f python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline --no-sweep > $O/c_ncu_e3.log 2>&1
echo "ncu e3 rc=$?"
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:pq_mma_filter -s 5 -c 1 -o $O/c_prof_last -f python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline --no-sweep > $O/c_ncu_last.log 2>&1
echo "ncu last rc=$?"
