"""torchrun --nnodes=1 --nproc-per-node R --master-addr 127.0.0.1 tests/workers/sharded_check.py
One process per GPU: ShardedIndexFlat (NCCL all-gather + merge kernel) with the threshold exchange over CUDA-IPC peer mailboxes,
against the oracle on rank 0.  Prints PASS/FAIL lines; exit code 1 on any failure."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import proqa_b200 as pq  # noqa: E402
from oracle import oracle  # noqa: E402
from proqa_b200.sharded import ShardedIndexFlat  # noqa: E402
from tests import data  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
fails = 0
for metric, nb, nq, k, layout in ((0, 400_000, 300, 100, "rows"), (1, 300_000, 64, 10, "rows"), (0, 400_000, 1500, 80, "rows"),
                                  (0, 500_000, 16, 2000, "rows"), (0, 400_000, 301, 100, "grid")):
    R = world if layout == "rows" else max(1, world // 2)
    xb, xq = data.corpus(nb), data.queries(nq)
    sh = ShardedIndexFlat(128, metric, device=local, row_shards=R)
    sh.add(xb)
    for rep in range(2):
        D, I = sh.search(xq, k)
    st = sh.local.last_stats
    if rank == 0:
        Dr, Ir = oracle.engine_spec(xq[:64], xb, k, metric)
        ok = np.array_equal(I[:64], Ir) and np.array_equal(D[:64].view(np.uint32), Dr.view(np.uint32))
        fails += 0 if ok else 1
        print(f"{'PASS' if ok else 'FAIL'} metric={metric} rows={nb} nq={nq} k={k} R={R} Q={world // R}: filter launches {st[3]}, "
              f"exchanges with every peer in time {st[9]}, repairs {st[8]}, fp32 re-runs {st[1]}", flush=True)
    del sh
dist.barrier()
dist.destroy_process_group()
sys.exit(1 if fails else 0)
