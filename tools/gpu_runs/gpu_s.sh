#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 200 -k "document_order" 2>&1 | tail -15
