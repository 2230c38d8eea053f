#!/bin/bash
# Wave-by-wave pacing on the 65,536-query batch (C3: 1024 CTAs, seven waves) A/B; flat-gather select on small batches; every GPU
# test; the default line.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q > $O/zd_pytest.log 2>&1
echo "gpu tests exit $?"; tail -4 $O/zd_pytest.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms", round(d["ms_per_step"],3), "roofline", d["roofline"]["bound"], round(d["roofline"]["frac"],4), "kern ms", round(d["roofline"]["kernel_ms_per_step"],3), d["roofline"].get("other_kernels_ms_per_step"),
          "parity", d["parity"]["ok"], "e2e ms", round(d["e2e"]["ms_per_step"],3), "launches", d["gpu_launches"], d.get("clocks"))
    for k_, v in (d.get("sweep") or {}).items(): print("   ", k_, {a: (round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a != "step_frac_of_hbm_note"})
except Exception as e:
    print("parse failed", sys.argv[1], e); print(open(sys.argv[1].replace(".json",".err")).read()[-2000:])
PY
}
for pace in 2 1 2 1; do
PROQA_B200_PACE=$pace timeout -s KILL 300 python bench.py --workload c3 --steps 5 --warmup 2 --no-cpu-baseline > $O/zd_c3_p${pace}.json 2> $O/zd_c3_p${pace}.err; show $O/zd_c3_p${pace}.json
done
for nq in 1 16 64 256; do
timeout -s KILL 200 python bench.py --workload s0 --nq $nq --steps 20 --warmup 3 --no-cpu-baseline > $O/zd_s0_nq$nq.json 2> $O/zd_s0_nq$nq.err; show $O/zd_s0_nq$nq.json
done
timeout -s KILL 200 python bench.py --workload s0 --nq 1 --k 1000 --steps 20 --warmup 3 --no-cpu-baseline > $O/zd_s0_nq1_k1000.json 2> $O/zd_s0_nq1_k1000.err; show $O/zd_s0_nq1_k1000.json
timeout -s KILL 300 python bench.py --workload c5 --rows 12500000 --steps 5 --warmup 2 --no-cpu-baseline > $O/zd_c5.json 2> $O/zd_c5.err; show $O/zd_c5.json
timeout -s KILL 200 python bench.py --workload trec --steps 10 --warmup 3 --no-cpu-baseline > $O/zd_trec.json 2> $O/zd_trec.err; show $O/zd_trec.json
timeout -s KILL 500 python bench.py --gpus 1 --steps 20 --warmup 3 > $O/zd_default.json 2> $O/zd_default.err; show $O/zd_default.json
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pq_ -c 80 --csv --log-file $O/zd_launches_s0_nq16.csv python bench.py --workload s0 --steps 1 --warmup 1 --no-cpu-baseline > $O/zd_ncu_s0.log 2>&1
python tools/launch_shares.py $O/zd_launches_s0_nq16.csv
