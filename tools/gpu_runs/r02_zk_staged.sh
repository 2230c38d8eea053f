#!/bin/bash
# Host-buffer searches with large query matrices / result arrays through the pinned double buffer: every GPU test, the k-means
# bench (2M points assigned through index.search from pageable memory), e2e of the C5 shard and the trec shape.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -q -x > $O/zk_pytest.log 2>&1
echo "gpu tests exit $?"; tail -3 $O/zk_pytest.log
timeout -s KILL 200 python tools/kmeans_bench.py > $O/zk_kmeans.log 2>&1; tr '\r' '\n' < $O/zk_kmeans.log | grep "final assignment\|total"
timeout -s KILL 200 python bench.py --workload c5 --rows 12500000 --steps 5 --warmup 2 --no-cpu-baseline > $O/zk_c5.json 2> $O/zk_c5.err
timeout -s KILL 100 python bench.py --workload trec --steps 10 --warmup 3 --no-cpu-baseline > $O/zk_trec.json 2> $O/zk_trec.err
python - <<'PY'
import json
for f in ("zk_c5", "zk_trec"):
    try:
        d = [json.loads(l) for l in open(f"gpurun_out/{f}.json").read().strip().splitlines() if l.startswith("{")][-1]
        print(f, "ms", round(d["ms_per_step"], 3), "parity", d["parity"]["ok"], "e2e ms", round(d["e2e"]["ms_per_step"], 2), "pinned e2e", d["e2e"].get("pinned_buffers_ms_per_step"))
    except Exception as e:
        print("parse failed", f, e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
