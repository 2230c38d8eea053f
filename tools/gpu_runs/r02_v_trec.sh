#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_largek.py tests/test_gpu_parity.py -m gpu -x -q > $O/v_pytest.log 2>&1
echo "gpu tests exit $?"; tail -3 $O/v_pytest.log
timeout -s KILL 300 python bench.py --workload trec --steps 10 --warmup 3 --no-cpu-baseline > $O/v_trec.json 2> $O/v_trec.err
timeout -s KILL 300 python bench.py --workload c5 --rows 12500000 --steps 5 --warmup 2 --no-cpu-baseline > $O/v_c5.json 2> $O/v_c5.err
python - <<'PY'
import json
for f in ("v_trec","v_c5"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "ms", round(d["ms_per_step"],3), "q/s", round(d["value"]), d["parity"]["ok"], "e2e", round(d["e2e"]["ms_per_step"],2))
    except Exception as e:
        print("parse failed", e); print(open(f"gpurun_out/{f}.err").read()[-2000:])
PY
