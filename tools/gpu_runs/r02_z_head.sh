#!/bin/bash
# Round 2, validation of HEAD on one GPU: every GPU test, smoke, the default bench line exactly as the driver runs it, the
# reference arm, the trec line, the launch list and one full ncu capture of the dominant launch (last epoch of a C2 search).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q > $O/z_pytest.log 2>&1
echo "gpu tests exit $?"; tail -4 $O/z_pytest.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/z_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/z_smoke.log
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
timeout -s KILL 500 python bench.py --gpus 1 --steps 20 --warmup 3 > $O/z_default.json 2> $O/z_default.err; show $O/z_default.json
timeout -s KILL 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/z_reference.json 2> $O/z_reference.err; tail -c 500 $O/z_reference.json; echo
timeout -s KILL 200 python bench.py --workload trec --steps 10 --warmup 3 --no-cpu-baseline > $O/z_trec.json 2> $O/z_trec.err; show $O/z_trec.json
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pq_ -c 60 --csv --log-file $O/z_launches_c2.csv python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline --no-sweep > $O/z_ncu_launch.log 2>&1
echo "ncu launch rc=$?"
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:pq_mma_filter -s 5 -c 1 -o $O/z_prof_last -f python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline --no-sweep > $O/z_ncu_full.log 2>&1
echo "ncu full rc=$?"
