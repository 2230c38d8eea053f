#!/bin/bash
# Round 2, final build on 8 GPUs: the default line as the driver launches it (both layouts), N = 4 as well, C3.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
run() { # name, nproc, extra args
  name=$1; n=$2; shift; shift
  timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus $n "$@" > $O/l_$name.json 2> $O/l_$name.err
  echo "$name rc=$?"
  python - $O/l_$name.json <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms", round(d["ms_per_step"],3), "q/s", d["value"] and round(d["value"]), d["config"]["parallelism"], "parity", d["parity"]["ok"], "e2e ms", round(d["e2e"]["ms_per_step"],3), d.get("clocks"))
    for k,v in d.get("layouts",{}).items(): print("  ", k, v)
except Exception as e:
    print("parse failed", e); print(open(sys.argv[1].replace(".json",".err")).read()[-3000:])
PY
}
run c2_g8 8 --steps 20 --warmup 5
run c2_g4 4 --steps 20 --warmup 5
run c3_g8 8 --workload c3 --steps 5 --warmup 2
