#!/bin/bash
# 1024-thread large-k finalize with warp-cooperative row reads; k = 1 finalize with a shorter dependent chain
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_largek.py tests/test_gpu_kmeans.py tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -m gpu -x -q > $O/y_pytest.log 2>&1
echo "gpu tests exit $?"; tail -3 $O/y_pytest.log
timeout -s KILL 300 python bench.py --workload trec --steps 10 --warmup 3 --no-cpu-baseline > $O/y_trec.json 2> $O/y_trec.err
for m in l2 ip; do
timeout -s KILL 300 python bench.py --workload c4 --metric $m --steps 5 --warmup 2 --no-cpu-baseline > $O/y_c4_$m.json 2> $O/y_c4_$m.err
done
python - <<'PY'
import json
for f in ("y_trec","y_c4_l2","y_c4_ip"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "ms", round(d["ms_per_step"],3), "value", round(d["value"]), "filter ms", round(d["roofline"]["kernel_ms_per_step"],2), "frac", round(d["roofline"]["frac"],3), d["parity"]["ok"], "e2e", round(d["e2e"]["ms_per_step"],2))
    except Exception as e:
        print("parse failed", f, e); print(open(f"gpurun_out/{f}.err").read()[-2000:])
PY
