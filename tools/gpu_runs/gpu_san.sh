#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout -s KILL 300 compute-sanitizer --tool $tool --error-exitcode 3 python tools/sanitize_small.py > gpurun_out/san2_$tool.log 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/san2_$tool.log
done
