#!/bin/bash
# Round 2, single-GPU validation of the final build: every GPU test, the default bench line (as the driver runs it), the
# reference arm, the other workloads' lines, the launch list.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q > $O/j_pytest.log 2>&1
echo "gpu tests exit $?"; tail -6 $O/j_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/j_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/j_smoke.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms", round(d["ms_per_step"],3), "q/s", d["value"] and round(d["value"]), "frac", round(d["roofline"]["frac"],4), "kern ms", round(d["roofline"]["kernel_ms_per_step"],3), d["roofline"].get("other_kernels_ms_per_step"),
          "launches", d["gpu_launches"], "parity", d["parity"]["ok"], "e2e ms", round(d["e2e"]["ms_per_step"],3), "cpu", d.get("cpu_baseline") and round(d["cpu_baseline"]["value"],2), d.get("clocks"))
    for k_, v in (d.get("sweep") or {}).items(): print("   ", k_, {a: (round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a != "step_frac_of_hbm_note"})
except Exception as e:
    print("parse failed", sys.argv[1], e); print(open(sys.argv[1].replace(".json",".err")).read()[-3000:])
PY
}
timeout -s KILL 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/j_default.json 2> $O/j_default.err; show $O/j_default.json
timeout -s KILL 400 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/j_reference.json 2> $O/j_reference.err; tail -c 600 $O/j_reference.json
timeout -s KILL 400 python bench.py --workload c3 --steps 5 --warmup 2 --no-cpu-baseline > $O/j_c3.json 2> $O/j_c3.err; show $O/j_c3.json
timeout -s KILL 300 python bench.py --workload trec --steps 10 --warmup 3 > $O/j_trec.json 2> $O/j_trec.err; show $O/j_trec.json
timeout -s KILL 300 python bench.py --workload c4 --steps 5 --warmup 2 --no-cpu-baseline > $O/j_c4_l2.json 2> $O/j_c4_l2.err; tail -c 900 $O/j_c4_l2.json; echo
timeout -s KILL 300 python bench.py --workload c4 --metric ip --steps 5 --warmup 2 --no-cpu-baseline > $O/j_c4_ip.json 2> $O/j_c4_ip.err; tail -c 300 $O/j_c4_ip.json; echo
timeout -s KILL 400 python bench.py --workload c5 --rows 12500000 --steps 5 --warmup 2 --no-cpu-baseline > $O/j_c5shard.json 2> $O/j_c5shard.err; show $O/j_c5shard.json
timeout -s KILL 300 python tools/kmeans_bench.py > $O/j_kmeans.log 2>&1; tail -5 $O/j_kmeans.log
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pq_ -c 60 --csv --log-file $O/j_launches_c2.csv python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline --no-sweep > $O/j_ncu_launch.log 2>&1
echo "ncu launch rc=$?"
