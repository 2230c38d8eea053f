#!/bin/bash
# Round 2, GPU call 5 (8 GPUs): the C2 line with both layouts, C3 (north_star's row-sharded configuration) with both.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
run() { # name, extra args
  name=$1; shift
  timeout -s KILL 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus 8 "$@" > $O/e_$name.json 2> $O/e_$name.err
  echo "$name rc=$?"
  python - $O/e_$name.json <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms", round(d["ms_per_step"],3), "q/s", d["value"] and round(d["value"]), d["config"]["parallelism"], "parity", d["parity"]["ok"], "e2e ms", round(d["e2e"]["ms_per_step"],3))
    for k,v in d.get("layouts",{}).items(): print("  ", k, v)
except Exception as e:
    print("parse failed", e); print(open(sys.argv[1].replace(".json",".err")).read()[-3000:])
PY
}
run c2 --steps 20 --warmup 3
run c3 --workload c3 --steps 5 --warmup 2
