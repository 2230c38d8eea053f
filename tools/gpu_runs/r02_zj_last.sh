#!/bin/bash
# Last call of round 2: every GPU test on the final build (batched fp32 scan included), the default line, C4 at full size.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -q > $O/zj_pytest.log 2>&1
echo "gpu tests exit $?"; tail -3 $O/zj_pytest.log
timeout -s KILL 400 python bench.py --gpus 1 --steps 20 --warmup 3 > $O/zj_default.json 2> $O/zj_default.err
for m in l2 ip; do
timeout -s KILL 200 python bench.py --workload c4 --metric $m --steps 5 --warmup 2 --no-cpu-baseline > $O/zj_c4_$m.json 2> $O/zj_c4_$m.err
done
python - <<'PY'
import json
for f in ("zj_default", "zj_c4_l2", "zj_c4_ip"):
    try:
        d = [json.loads(l) for l in open(f"gpurun_out/{f}.json").read().strip().splitlines() if l.startswith("{")][-1]
        print(f, "ms", round(d["ms_per_step"], 3), "value", round(d["value"]), "kern ms", round(d["roofline"]["kernel_ms_per_step"], 2), "frac", round(d["roofline"]["frac"], 4),
              "parity", d["parity"]["ok"], "reruns/step", d["parity"].get("fp32_rerun_queries_per_step"), "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"], 2), d.get("clocks"))
        for k_, v in (d.get("sweep") or {}).items(): print("   ", k_, v.get("ms") and round(v["ms"], 4), v.get("parity_ok"))
    except Exception as e:
        print("parse failed", f, e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
