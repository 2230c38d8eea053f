#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_xchg.py -m gpu -q > $O/g_xchg_tests.log 2>&1
echo "xchg tests exit $?"; tail -12 $O/g_xchg_tests.log
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 tests/workers/sharded_check.py > $O/g_sharded_check.log 2>&1
echo "sharded check rc=$?"; grep -E "PASS|FAIL|Error|error" $O/g_sharded_check.log | head -20
timeout -s KILL 900 python -m pytest tests/test_gpu_baseline_sizes.py tests/test_gpu_sharded.py -m gpu -q > $O/g_pytest.log 2>&1
echo "baseline-size + sharded tests exit $?"; tail -12 $O/g_pytest.log
