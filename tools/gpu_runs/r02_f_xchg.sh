#!/bin/bash
# Round 2, GPU call 6 (2 GPUs): peer-memory exchange (pq_xchg) single-device tests, then two ranks under torchrun; mid-k (C5 shard) line.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_xchg.py -m gpu -x -q > $O/f_xchg_tests.log 2>&1
echo "xchg tests exit $?"; tail -12 $O/f_xchg_tests.log
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 tests/workers/sharded_check.py > $O/f_sharded_check.log 2>&1
echo "sharded check rc=$?"; grep -E "PASS|FAIL|Error|error" $O/f_sharded_check.log | head -20
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > $O/f_pytest.log 2>&1
echo "all gpu tests exit $?"; tail -6 $O/f_pytest.log
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29623 bench.py --gpus 2 --steps 20 --warmup 3 > $O/f_c2_g2.json 2> $O/f_c2_g2.err
echo "bench g2 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/f_c2_g2.json").read().strip().splitlines()[-1])
    print("c2 g2", d["ms_per_step"], d["value"], d["config"]["parallelism"], d["parity"]["ok"], "e2e", d["e2e"]["ms_per_step"])
    for k,v in d["layouts"].items(): print("  ", k, v)
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/f_c2_g2.err").read()[-3000:])
PY
# C5 shard on one GPU: 8192 queries x 12.5M rows, k = 1000 (mid-k scheme)
timeout -s KILL 400 python bench.py --workload c5 --rows 12500000 --steps 5 --warmup 2 --no-cpu-baseline > $O/f_c5shard.json 2> $O/f_c5shard.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/f_c5shard.json").read().strip().splitlines()[-1])
    print("c5 shard", d["ms_per_step"], d["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_step"], d["parity"], d["gpu_launches"])
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/f_c5shard.err").read()[-3000:])
PY
