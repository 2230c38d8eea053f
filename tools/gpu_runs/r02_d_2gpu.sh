#!/bin/bash
# Round 2, GPU call 4 (2 GPUs): row shards on two devices — one process (pq_multi) and one process per GPU (torchrun, CUDA IPC
# mailboxes, NCCL all-gather + merge kernel); then the C2 line on 2 GPUs with both layouts.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/d_smi.log 2>&1; cat $O/d_smi.log
timeout -s KILL 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sharded.py tests/test_gpu_sharded_kmeans.py tests/test_gpu_reference_script.py -m gpu -x -q > $O/d_pytest.log 2>&1
echo "2-gpu tests exit $?"; tail -15 $O/d_pytest.log
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/workers/sharded_check.py > $O/d_sharded_check.log 2>&1
echo "sharded check rc=$?"; grep -E "PASS|FAIL|Error|error" $O/d_sharded_check.log | head -20
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tests/workers/sharded_kmeans.py > $O/d_sharded_kmeans.log 2>&1
echo "sharded kmeans rc=$?"; tail -3 $O/d_sharded_kmeans.log
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 20 --warmup 3 > $O/d_c2_g2.json 2> $O/d_c2_g2.err
echo "bench g2 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/d_c2_g2.json").read().strip().splitlines()[-1])
    print("c2 g2", d["ms_per_step"], d["value"], d["config"]["parallelism"], d["parity"])
    for k,v in d["layouts"].items(): print("  ", k, v)
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/d_c2_g2.err").read()[-3000:])
PY
