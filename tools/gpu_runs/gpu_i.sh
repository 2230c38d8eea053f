#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests -m gpu -q -x --timeout 150 2>&1 | tail -8 > gpurun_out/i_pytest.log
cat gpurun_out/i_pytest.log
