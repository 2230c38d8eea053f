#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
time (timeout -s KILL 400 python bench.py > gpurun_out/n_default.json 2> gpurun_out/n_default.err) || { echo "default bench failed"; tail -5 gpurun_out/n_default.err; }
cat gpurun_out/n_default.json
