#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests -m gpu -q -x --timeout 120 2>&1 | tail -25 > gpurun_out/i_pytest.log
cat gpurun_out/i_pytest.log
