#!/usr/bin/env python
"""Per-kernel shares of ONE search step from an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`).

    python tools/launch_shares.py gpurun_out/z_launches_c2.csv [step_index]

A step starts at every pq_mma_init_state_kernel (one per search on the tensor tier); step_index picks which one (default: the
last complete one).  ncu serialises the launches and runs them cold-cache: compare SHARES with the bench line, not absolutes."""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10 and r[0].isdigit()]
    launches = [(re.sub(r"^void ", "", re.sub(r"\(.*", "", r[4])), float(r[-1].replace(",", "")) / 1e3) for r in rows]  # (name, us)
    starts = [i for i, (n, _) in enumerate(launches) if n.startswith("pq_mma_init_state_kernel")]
    if not starts:
        raise SystemExit("no pq_mma_init_state_kernel in the list: not a tensor-tier search")
    bounds = [(s, e) for s, e in zip(starts, starts[1:])]
    if len(sys.argv) > 2:
        s, e = bounds[int(sys.argv[2])]
    else:
        s, e = bounds[-1]
    step = [l for l in launches[s:e] if not l[0].startswith("pq_prep_rows_kernel")]   # (the next search's query preparation)
    prep = [l for l in launches[max(0, s - 1):s] if l[0].startswith("pq_prep_rows_kernel")]
    step = prep + step
    total = sum(us for _, us in step)
    by = OrderedDict()
    for n, us in step:
        by.setdefault(n, []).append(us)
    print(f"# launches in the step: {len(step)}; total {total / 1e3:.3f} ms under ncu")
    for n, v in by.items():
        each = ", ".join(f"{x:.1f}" for x in v)
        print(f"{n:<52} launches={len(v):3d} total_ms={sum(v) / 1e3:9.3f} share_of_step={sum(v) / total:.3f} each_us=[{each}]")


if __name__ == "__main__":
    main()
