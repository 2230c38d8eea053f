#!/bin/bash
# Final build of round 2 on one GPU: every GPU test, smoke, the default line and the reference arm as the driver runs them, the
# other workloads' lines, the launch list of a C2 search.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q > $O/zf_pytest.log 2>&1
echo "gpu tests exit $?"; tail -4 $O/zf_pytest.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/zf_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/zf_smoke.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms", round(d["ms_per_step"],3), "roofline", d["roofline"]["bound"], round(d["roofline"]["frac"],4), "kern ms", round(d["roofline"]["kernel_ms_per_step"],3), d["roofline"].get("other_kernels_ms_per_step"),
          "parity", d["parity"]["ok"], "e2e ms", round(d["e2e"]["ms_per_step"],3), "launches", d["gpu_launches"], "cpu", d.get("cpu_baseline") and round(d["cpu_baseline"]["value"],2), d.get("clocks"))
    for k_, v in (d.get("sweep") or {}).items(): print("   ", k_, {a: (round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a != "step_frac_of_hbm_note"})
except Exception as e:
    print("parse failed", sys.argv[1], e); print(open(sys.argv[1].replace(".json",".err")).read()[-2000:])
PY
}
# (reference arm: see the previous call)
timeout -s KILL 500 python bench.py --gpus 1 --steps 20 --warmup 3 > $O/zf_default.json 2> $O/zf_default.err; show $O/zf_default.json
timeout -s KILL 200 python bench.py --workload trec --steps 10 --warmup 3 --no-cpu-baseline > $O/zf_trec.json 2> $O/zf_trec.err; show $O/zf_trec.json
timeout -s KILL 300 python bench.py --workload c5 --rows 12500000 --steps 5 --warmup 2 --no-cpu-baseline > $O/zf_c5.json 2> $O/zf_c5.err; show $O/zf_c5.json
timeout -s KILL 200 python bench.py --workload s0 --nq 1 --k 5000 --steps 20 --warmup 3 --no-cpu-baseline > $O/zf_s0_nq1_k5000.json 2> $O/zf_s0_nq1_k5000.err; show $O/zf_s0_nq1_k5000.json
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pq_ -c 60 --csv --log-file $O/zf_launches_c2.csv python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline --no-sweep > $O/zf_ncu_launch.log 2>&1
python tools/launch_shares.py $O/zf_launches_c2.csv
