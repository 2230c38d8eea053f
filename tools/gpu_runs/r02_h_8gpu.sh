#!/bin/bash
# Round 2, GPU call 8 (8 GPUs): C2 / C3 / C5 with the peer-memory exchange and the corrected shard epochs, both layouts each.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
run() { # name, extra args
  name=$1; shift
  timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus 8 "$@" > $O/h_$name.json 2> $O/h_$name.err
  echo "$name rc=$?"
  python - $O/h_$name.json <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms", round(d["ms_per_step"],3), "q/s", d["value"] and round(d["value"]), d["config"]["parallelism"], "parity", d["parity"]["ok"], "e2e ms", round(d["e2e"]["ms_per_step"],3), d.get("clocks"))
    for k,v in d.get("layouts",{}).items(): print("  ", k, v)
except Exception as e:
    print("parse failed", e); print(open(sys.argv[1].replace(".json",".err")).read()[-3000:])
PY
}
run c2 --steps 20 --warmup 3
run c3 --workload c3 --steps 5 --warmup 2
run c5 --workload c5 --steps 5 --warmup 2
