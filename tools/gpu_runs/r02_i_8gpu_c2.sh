#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 8 --steps 20 --warmup 3 --row-shards rows > $O/i_c2.json 2> $O/i_c2.err
python - $O/i_c2.json <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("ms", round(d["ms_per_step"],3), d["config"]["parallelism"], "parity", d["parity"]["ok"], "e2e ms", round(d["e2e"]["ms_per_step"],3), d["e2e"].get("pinned_buffers_ms_per_step"))
    print(d["roofline"]); print(d["multi_gpu_phases"])
except Exception as e:
    print("parse failed", e); print(open(sys.argv[1].replace(".json",".err")).read()[-3000:])
PY
