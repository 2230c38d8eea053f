#!/bin/bash
# A/B of the filter kernel's epilogue fast path (one vote per accumulator): new build vs the library of the previous commit
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py tests/test_gpu_largek.py -m gpu -x -q > $O/w_pytest.log 2>&1
echo "gpu tests exit $?"; tail -3 $O/w_pytest.log
cp proqa_b200/libproqa_b200.so /tmp/new.so
for i in 1 2; do
  cp /tmp/new.so proqa_b200/libproqa_b200.so
  timeout -s KILL 200 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline --no-sweep > $O/w_c2_new_$i.json 2> $O/w_c2_new_$i.err
  cp tools/ab/libproqa_b200_old.so proqa_b200/libproqa_b200.so
  timeout -s KILL 200 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline --no-sweep > $O/w_c2_old_$i.json 2> $O/w_c2_old_$i.err
done
cp /tmp/new.so proqa_b200/libproqa_b200.so
python - <<'PY'
import json
for f in ("w_c2_new_1","w_c2_old_1","w_c2_new_2","w_c2_old_2"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        r=d["roofline"]
        print(f, "ms", round(d["ms_per_step"],3), "kernel_ms", round(r["kernel_ms_per_step"],3), "frac", round(r["frac"],4), "clk", d["clocks"]["sm_mhz"], d["parity"]["ok"], "e2e", round(d["e2e"]["ms_per_step"],2))
    except Exception as e:
        print("parse failed", f, e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
