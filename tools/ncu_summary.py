#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): key raw metrics + top stall reasons + hottest source lines."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct", "sm__inst_executed_pipe_fma.avg.pct",
        "sm__inst_executed_pipe_fmaheavy", "sm__inst_executed_pipe_lsu.avg.pct", "sm__warps_active.avg.pct", "launch__registers_per_thread ",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum ", "sm__throughput.avg.pct", "smsp__issue_active.avg.pct",
        "smsp__cycles_active.avg ", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ", "lts__t_bytes.sum ", "sm__cycles_elapsed.avg ",
        "smsp__inst_executed_pipe_uniform", "sm__inst_executed_pipe_tmem", "smsp__average_warp", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__pcsamp_warps_issue_stalled", "sm__pipe_shared_cycles_active"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== kernel:", r[hdr.index("Kernel Name")][:80], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
        stalls = []
        for i, h in enumerate(hdr):
            if any(k.strip() in h for k in KEYS) and "pcsamp" not in h:
                print(f"   {h} [{units[i]}] = {r[i]}")
            if "smsp__pcsamp_warps_issue_stalled" in h and not h.endswith("_not_issued"):
                try:
                    stalls.append((float(r[i].replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(s for s, _ in stalls) or 1
        print("   stall samples:", ", ".join(f"{n}={s / tot:.1%}" for s, n in sorted(stalls, reverse=True)[:10]))


def source(rep, top=25):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    if not hi:
        print("   (no source page)")
        return
    hdr = rows[hi[0]]
    si, ci, ei = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for r in rows[hi[0] + 1:]:
        try:
            n = float(r[ci].replace(",", "") or 0)
            tops = sorted(((float(r[i] or 0), h) for i, h in stall_cols), reverse=True)[:2]
            data.append((n, f"{r[si].strip()[:70]:70s} exec={r[ei]:>10s} " + " ".join(f"{h[6:]}={v:.0f}" for v, h in tops if v > 0)))
        except (ValueError, IndexError):
            pass
    tot = sum(d for d, _ in data) or 1
    print(f"   hottest SASS by stall samples (total {tot:.0f}):")
    for d, s in sorted(data, reverse=True)[:top]:
        print(f"     {d / tot:6.1%}  {s}")


if __name__ == "__main__":
    raw(sys.argv[1])
    if len(sys.argv) > 2:
        source(sys.argv[1], int(sys.argv[2]))
