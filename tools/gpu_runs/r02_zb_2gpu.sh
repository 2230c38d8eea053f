#!/bin/bash
# Final build on 2 GPUs: multi-GPU tests (one process and one process per GPU), the sharded check, the C2 line with both layouts.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L | head -3
timeout -s KILL 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sharded.py tests/test_gpu_sharded_kmeans.py tests/test_gpu_reference_script.py tests/test_gpu_xchg.py -m gpu -q > $O/zb_pytest.log 2>&1
echo "2-gpu tests exit $?"; tail -5 $O/zb_pytest.log
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/workers/sharded_check.py > $O/zb_sharded_check.log 2>&1
echo "sharded check rc=$?"; grep -E "PASS|FAIL|Error|error" $O/zb_sharded_check.log | head -20
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 20 --warmup 3 > $O/zb_c2_g2.json 2> $O/zb_c2_g2.err
echo "bench g2 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/zb_c2_g2.json").read().strip().splitlines()[-1])
    print("c2 g2", d["ms_per_step"], d["value"], d["config"]["parallelism"], d["parity"], d.get("clocks"))
    for k,v in d["layouts"].items(): print("  ", k, v)
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/zb_c2_g2.err").read()[-3000:])
PY
