#!/usr/bin/env python
"""bench.py — exact top-k MIPS throughput on B200 (BASELINE.json metric) with roofline and CPU baseline.

    python bench.py --gpus 1 --steps K --warmup W            # our arm, N = 1
    torchrun --nproc-per-node N ... bench.py --gpus N ...     # our arm, corpus row-sharded over N ranks
    python bench.py --impl reference ...                      # the reference's CPU algorithm on host cores

One "step" = one search of the whole query batch against the whole (resident) corpus.  Default workload is
BASELINE.json configs[1] ("c2"): 3610 queries x 21M x 128 fp32 corpus, k = 100.  Prints ONE JSON line.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "c1": dict(nq=2032, rows=1_000_000, k=80, desc="eval_retrieval.py shape: 2032 queries x 1M x 128 fp32, k=80"),
    "c2": dict(nq=3610, rows=21_000_000, k=100, desc="NQ-scale search: 3610 queries x 21M x 128 fp32 corpus, k=100"),
    "c3": dict(nq=65536, rows=21_000_000, k=80, desc="large batch: 65536 queries x 21M x 128 fp32, k=80"),
    "s0": dict(nq=16, rows=21_000_000, k=80, desc="small-batch sweep point: 16 queries x 21M x 128 fp32, k=80 (HBM-bound)"),
    "c4": dict(nq=21_000_000, rows=10_000, k=1, kind="kmeans",
               desc="group_paras.py k-means assignment: 21M x 128 points x 10,000 centroids, k=1 (points are the query side)"),
    "trec": dict(nq=256, rows=8_841_823, k=10000,
                 desc="trec_process.py:76 shape: a 256-query sample of index.search(xq, 10000) over 8.8M x 128 fp32 (MS MARCO passages)"),
    "c5": dict(nq=8192, rows=100_000_000, k=1000, desc="scale-out: 8192 queries x 100M x 128 fp32, k=1000 (needs 8 GPUs for the full corpus)"),
}
CHUNK = 1_000_000  # rows per generated chunk; chunk c is seeded with 1234 + c so the corpus does not depend on N


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm_gbs=float(p["hbm_gbs"]), bf16_tflops=float(p["bf16_tflops"]),
                    bf16_tflops_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), source="measured")
    except Exception:
        return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


def host_queries(nq):
    return np.random.default_rng(4321).standard_normal((nq, 128), dtype=np.float32)


def host_corpus_sample(rows):
    return np.random.default_rng(1234).standard_normal((rows, 128), dtype=np.float32)


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the FAISS-1.6.3 restatement (oracle) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_search_time(xq, k, rows, reps=1):
    from oracle import oracle
    xb = host_corpus_sample(rows)
    ix = oracle.IndexFlatIP(128)
    ix.add(xb)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        ix.search(xq, k)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best


def cpu_baseline(wl, budget_s=12.0):
    """Bounded sample of the same workload: all queries, a prefix of the corpus; linear extrapolation in N."""
    xq = host_queries(wl["nq"])
    probe_rows = 20_000
    t_probe = cpu_search_time(xq, wl["k"], probe_rows)
    rows = int(min(wl["rows"], max(probe_rows, probe_rows * budget_s / max(t_probe, 1e-6))))
    rows = min(rows, 2_000_000)
    t = cpu_search_time(xq, wl["k"], rows)
    full_t = t * wl["rows"] / rows
    return {"value": wl["nq"] / full_t, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"all {wl['nq']} queries x first {rows} of {wl['rows']} rows in {t:.2f}s, extrapolated linearly in rows; "
                      "FAISS-1.6.3 restatement (OpenBLAS sgemm 4096x1024 blocks + per-query heap), not FAISS"}


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if wl.get("kind") == "kmeans":
        return run_reference_kmeans(args, wl)
    xq = host_queries(wl["nq"])
    total = args.steps + args.warmup
    per_step_budget = max(2.0, min(20.0, 150.0 / max(total, 1)))
    probe_rows = 20_000
    t_probe = cpu_search_time(xq, wl["k"], probe_rows)
    rows = int(min(wl["rows"], 2_000_000, max(probe_rows, probe_rows * per_step_budget / max(t_probe, 1e-6))))
    from oracle import oracle
    ix = oracle.IndexFlatIP(128)
    ix.add(host_corpus_sample(rows))
    for _ in range(args.warmup):
        ix.search(xq, wl["k"])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ix.search(xq, wl["k"])
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    full_t = dt * wl["rows"] / rows
    value = wl["nq"] / full_t
    sample = (f"each step: all {wl['nq']} queries x first {rows} of {wl['rows']} rows ({dt:.2f}s), extrapolated linearly in rows; "
              "FAISS-1.6.3 restatement (OpenBLAS sgemm + heap), not FAISS (faiss-cpu is not installable here)")
    out = {"impl": "reference", "metric": "queries_per_sec", "value": value, "unit": "queries/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": full_t * 1e3, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, wl),
           "corpus_gbs": wl["rows"] * 512 / full_t / 1e9,
           "cpu_baseline": {"value": value, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def run_reference_kmeans(args, wl):
    """C4 on the host cores: a bounded sample of points against all centroids per step (IndexFlatL2/IP.search(x, 1))."""
    from oracle import oracle
    cents = np.random.default_rng(777).standard_normal((wl["rows"], 128), dtype=np.float32)
    fo = oracle.FaissFlatOracle(128, 1 if args.metric == "l2" else 0)
    fo.add(cents)
    total = args.steps + args.warmup
    per_step_budget = max(1.0, min(10.0, 120.0 / max(total, 1)))
    probe = host_queries(20_000)
    t0 = time.perf_counter()
    fo.search(probe, 1)
    t_probe = time.perf_counter() - t0
    ns = int(min(wl["nq"], 2_000_000, max(20_000, 20_000 * per_step_budget / max(t_probe, 1e-6))))
    xs = host_queries(ns)
    for _ in range(args.warmup):
        fo.search(xs, 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fo.search(xs, 1)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    value = ns / dt
    sample = (f"each step: {ns} of {wl['nq']} points x all {wl['rows']} centroids ({dt:.2f}s); FAISS-1.6.3 restatement "
              "(OpenBLAS sgemm + heap), not FAISS (faiss-cpu is not installable here)")
    out = {"impl": "reference", "metric": "queries_per_sec", "value": value, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": wl["nq"] / value * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": wl["desc"], "name": args.workload, "points": wl["nq"], "centroids": wl["rows"], "d": 128, "k": 1, "metric": args.metric},
           "corpus_gbs": value * 512 / 1e9,
           "cpu_baseline": {"value": value, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def workload_config(args, wl):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from proqa_b200.sharded import auto_row_shards
    rs = getattr(args, "row_shards", None)
    R = auto_row_shards(world, wl["rows"]) if rs in (None, "auto") else (world if rs == "rows" else int(rs))
    return {"workload": wl["desc"], "name": args.workload, "nq": wl["nq"], "rows": wl["rows"], "d": 128, "k": wl["k"],
            "parallelism": f"{R} row shards x {world // R} query groups" if world > 1 else "1 GPU",
            "l2_flush": "inputs larger than L2 (bf16 corpus copy %.1f GB per step)" % (wl["rows"] * 256 / 1e9)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def build_shard(index, lo, hi, dev, sharded=None, n_global=None):
    """Generate rows [lo, hi) of the synthetic corpus on the device, chunk by chunk, and append them."""
    import torch
    first = True
    c0, c1 = lo // CHUNK, (hi + CHUNK - 1) // CHUNK
    for c in range(c0, c1):
        g = torch.Generator(device=dev)
        g.manual_seed(1234 + c)
        rows = min(CHUNK, n_global - c * CHUNK)
        x = torch.randn((rows, 128), generator=g, device=dev, dtype=torch.float32)
        a, b = max(lo, c * CHUNK) - c * CHUNK, min(hi, c * CHUNK + rows) - c * CHUNK
        part = x[a:b].contiguous()
        torch.cuda.synchronize()
        if first and sharded is not None:
            sharded._local.set_id_base(lo)
        first = False
        index.add_device(part.data_ptr(), part.shape[0])
        del x, part


def truth_topk_fp64(xq_dev, lo, hi, n_global, k, dev, ids=None):
    """fp64 ground truth for a query slice over rows [lo, hi) — checker only (torch matmul in float64).  With `ids` [nq, k] also
    returns the fp64 score of every id that falls into [lo, hi) (0 elsewhere) — what the engine's answer is REALLY worth."""
    import torch
    q = xq_dev.double()
    best_d = torch.full((q.shape[0], k), -float("inf"), dtype=torch.float64, device=dev)
    best_i = torch.full((q.shape[0], k), -1, dtype=torch.int64, device=dev)
    own = torch.zeros(ids.shape, dtype=torch.float64, device=dev) if ids is not None else None
    for c in range(lo // CHUNK, (hi + CHUNK - 1) // CHUNK):
        g = torch.Generator(device=dev)
        g.manual_seed(1234 + c)
        rows = min(CHUNK, n_global - c * CHUNK)
        x = torch.randn((rows, 128), generator=g, device=dev, dtype=torch.float32)
        a, b = max(lo, c * CHUNK) - c * CHUNK, min(hi, c * CHUNK + rows) - c * CHUNK
        s = q @ x[a:b].double().T
        d, i = torch.topk(s, min(k, s.shape[1]), dim=1)
        i = i + (c * CHUNK + a)
        cat_d, cat_i = torch.cat([best_d, d], 1), torch.cat([best_i, i], 1)
        top = torch.topk(cat_d, k, dim=1)
        best_d, best_i = top.values, torch.gather(cat_i, 1, top.indices)
        if ids is not None:
            inside = (ids >= c * CHUNK + a) & (ids < c * CHUNK + b)
            loc = torch.where(inside, ids - (c * CHUNK + a), torch.zeros_like(ids))
            own += torch.where(inside, torch.gather(s, 1, loc), torch.zeros_like(own))
        del x, s
    return best_d, best_i, own


def parity_gate(D, I, Dt, It, own, rtol=1e-4):
    """north_star tolerance, checked on what was RETURNED: (1) every reported score is within rtol of the fp64 score of the id
    reported with it; (2) the fp64 scores of the returned ids are, rank by rank, within rtol of the true top-k scores (so the
    set is the true top-k up to near-ties and the order is best-first up to near-ties); (3) where an id differs from the fp64
    ranking, the gap between the two candidates' fp64 scores is reported (largest one: `max_near_tie_gap_rel`)."""
    import torch
    D64 = D.double()
    scale = torch.maximum(Dt.abs(), torch.full_like(Dt, 1e-3))
    tol = rtol * scale + 1e-6
    score_ok = bool(((D64 - own).abs() <= tol).all())
    rank_ok = bool(((own - Dt).abs() <= tol).all())
    neq = I != It
    frac_equal = 1.0 - float(neq.double().mean())
    gap = float((((own - Dt).abs() / scale) * neq.double()).max()) if bool(neq.any()) else 0.0
    return score_ok and rank_ok, {"ids_equal_frac": frac_equal, "max_near_tie_gap_rel": gap, "scores_match_returned_ids": score_ok,
                                   "returned_ids_are_topk_within_tol": rank_ok}


def device_roofline(ix, run, flops_per_step, bytes_per_step, t_dev_hint, traffic_name=None):
    """Dominant-kernel time from CUDA events around it on its launching stream (pq_index_set_profile), against the measured peaks."""
    ix.set_profile(True)
    kern_us, other_us = [], []
    for _ in range(3):
        run()
        kern_us.append(ix.last_stats[7])
        other_us.append(list(ix.last_stats[11:14]))
    ix.set_profile(False)
    st = ix.last_stats
    kernel_s = float(np.mean(kern_us)) * 1e-6
    other = {"epoch_select_ms": float(np.mean([o[0] for o in other_us])) / 1e3, "threshold_fold_ms": float(np.mean([o[1] for o in other_us])) / 1e3,
             "rescore_ms": float(np.mean([o[2] for o in other_us])) / 1e3}
    peaks = load_peaks()
    if st[3] > 0:
        long_run = t_dev_hint > 1.0
        peak = peaks["bf16_tflops_sustained"] if long_run else peaks["bf16_tflops"]
        achieved = flops_per_step / kernel_s / 1e12 if kernel_s > 0 else 0.0
        tr = load_traffic("pq_mma_filter_kernel") if traffic_name else None
        return {"bound": "tensor", "kernel": "pq_mma_filter_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": tr,
                "traffic_note": "archived ncu --set full capture of the last epoch of a C2 search (profiles/), not measured in this run" if tr else None,
                "peak_source": f"{peaks['source']} cuBLAS bf16 ({'sustained' if long_run else 'burst'})",
                "frac_of_burst": achieved / peaks["bf16_tflops"], "frac_of_sustained": achieved / peaks["bf16_tflops_sustained"],
                "algorithmic": "256 flop per (query,row) score",
                "launches_per_step": int(st[3]), "kernel_ms_per_step": kernel_s * 1e3, "other_kernels_ms_per_step": other}
    nbytes = bytes_per_step * max(1, st[2])
    achieved = nbytes / kernel_s / 1e9 if kernel_s > 0 else 0.0
    tr = load_traffic("pq_ffma_scan_kernel") if traffic_name else None
    return {"bound": "hbm", "kernel": "pq_ffma_scan_kernel", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / peaks["hbm_gbs"], "traffic": tr,
            "traffic_note": "archived ncu --set full capture of one scan over 21M rows (profiles/), not measured in this run" if tr else None,
            "peak_source": f"{peaks['source']} copy bandwidth", "algorithmic": "512 B of corpus per row per pass",
            "launches_per_step": int(st[2]), "kernel_ms_per_step": kernel_s * 1e3}


def timed(stream, fn, steps, warmup):
    """ms per call: CUDA events on the launching stream around `steps` calls, after `warmup` calls."""
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def sweep_lines(args, ix, dev, stream, N, local_rank):
    """Driver-visible lines for the other tiers, on the resident corpus where possible (single GPU): S0 small-batch sweep
    (north_star: >= 70 % of the HBM roofline), C1 (BASELINE configs[0]) and a slice of C4 (k-means assignment).  A few steps each."""
    import torch
    import proqa_b200 as pq
    out = {}
    peaks = load_peaks()
    k = 80
    for nq in (1, 4, 8, 16, 64, 256):
        q = torch.from_numpy(host_queries(nq)).to(dev)
        D = torch.empty((nq, k), dtype=torch.float32, device=dev)
        I = torch.empty((nq, k), dtype=torch.int64, device=dev)
        run = lambda: ix.search_device(q.data_ptr(), nq, k, D.data_ptr(), I.data_ptr())  # noqa: E731
        ms = timed(stream, run, 10, 3)
        Dt, It, own = truth_topk_fp64(q, 0, N, N, k, dev, ids=I)
        ok, _ = parity_gate(D, I, Dt, It, own)
        rf = device_roofline(ix, run, 2.0 * nq * N * 128, 512.0 * N, 0.0)
        # what the batch costs against streaming the corpus once: fp32 rows for the scan, the bf16 copy for the tensor tier
        hbm_bytes = N * (512.0 if rf["bound"] == "hbm" else 256.0)
        out[f"s0_nq{nq}"] = {"nq": nq, "rows": N, "k": k, "ms": ms, "queries_per_s": nq / ms * 1e3, "corpus_gbs_fp32_equiv": N * 512 / ms / 1e6,
                             "tier": rf["bound"], "kernel_frac_of_its_roofline": rf["frac"],
                             "step_frac_of_hbm": hbm_bytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                             "step_frac_of_hbm_note": "bytes the tier must stream once (fp32 rows: scan; bf16 copy: tensor tier) / whole-step time / measured copy bandwidth",
                             "parity_ok": ok}
        if nq <= 4:
            # AUTO answers these from the bf16 copy on a corpus this large; the exact fp32 scan (north_star (2): the HBM-bound FFMA tier,
            # what smaller corpora and certificate failures get) is timed beside it
            ix.set_tier("fp32")
            ms = timed(stream, run, 10, 3)
            Dt, It, own = truth_topk_fp64(q, 0, N, N, k, dev, ids=I)
            ok, _ = parity_gate(D, I, Dt, It, own)
            rf = device_roofline(ix, run, 2.0 * nq * N * 128, 512.0 * N, 0.0)
            ix.set_tier("auto")
            out[f"s0_nq{nq}_fp32_scan"] = {"nq": nq, "rows": N, "k": k, "ms": ms, "queries_per_s": nq / ms * 1e3,
                                           "corpus_gbs_fp32_equiv": N * 512 / ms / 1e6, "tier": rf["bound"],
                                           "kernel_frac_of_its_roofline": rf["frac"],
                                           "step_frac_of_hbm": N * 512.0 / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "parity_ok": ok}
    # C1: eval_retrieval.py shape, its own 1M-row index
    wl = WORKLOADS["c1"]
    c1 = pq.IndexFlatIP(128, local_rank)
    build_shard(c1, 0, wl["rows"], dev, n_global=wl["rows"])
    c1.set_stream(stream.cuda_stream)
    q = torch.from_numpy(host_queries(wl["nq"])).to(dev)
    D = torch.empty((wl["nq"], wl["k"]), dtype=torch.float32, device=dev)
    I = torch.empty((wl["nq"], wl["k"]), dtype=torch.int64, device=dev)
    run = lambda: c1.search_device(q.data_ptr(), wl["nq"], wl["k"], D.data_ptr(), I.data_ptr())  # noqa: E731
    ms = timed(stream, run, 10, 3)
    Dt, It, own = truth_topk_fp64(q[:256], 0, wl["rows"], wl["rows"], wl["k"], dev, ids=I[:256])
    ok, _ = parity_gate(D[:256], I[:256], Dt, It, own)
    rf = device_roofline(c1, run, 2.0 * wl["nq"] * wl["rows"] * 128, 512.0 * wl["rows"], 0.0)
    out["c1"] = {"nq": wl["nq"], "rows": wl["rows"], "k": wl["k"], "ms": ms, "queries_per_s": wl["nq"] / ms * 1e3,
                 "kernel_frac_of_its_roofline": rf["frac"], "tier": rf["bound"], "parity_ok": ok}
    del c1, q, D, I
    # C4 slice: 2M points against 10,000 centroids, k = 1, both metrics
    g = torch.Generator(device=dev)
    g.manual_seed(777)
    cents = torch.randn((10_000, 128), generator=g, device=dev, dtype=torch.float32)
    g.manual_seed(4321)
    pts = torch.randn((2_000_000, 128), generator=g, device=dev, dtype=torch.float32)
    for name, metric in (("c4_l2_2m", pq.METRIC_L2), ("c4_ip_2m", pq.METRIC_INNER_PRODUCT)):
        km = pq.IndexFlat(128, metric, local_rank)
        torch.cuda.synchronize()
        km.add_device(cents.data_ptr(), 10_000)
        km.set_stream(stream.cuda_stream)
        D = torch.empty((len(pts), 1), dtype=torch.float32, device=dev)
        I = torch.empty((len(pts), 1), dtype=torch.int64, device=dev)
        run = lambda: km.search_device(pts.data_ptr(), len(pts), 1, D.data_ptr(), I.data_ptr())  # noqa: E731
        ms = timed(stream, run, 5, 2)
        p64, c64 = pts[:4096].double(), cents.double()
        S = p64 @ c64.T
        if metric == pq.METRIC_L2:
            best, arg = ((p64 * p64).sum(1, keepdim=True) + (c64 * c64).sum(1)[None, :] - 2.0 * S).min(1)
        else:
            best, arg = S.max(1)
        scale = torch.maximum(best.abs(), torch.full_like(best, 1e-3))
        ok = bool(((D[:4096, 0].double() - best).abs() <= 1e-4 * scale + 1e-6).all()) and float((I[:4096, 0] == arg).double().mean()) > 0.995
        rf = device_roofline(km, run, 2.0 * len(pts) * 10_000 * 128, 0.0, 0.0)
        out[name] = {"points": len(pts), "centroids": 10_000, "k": 1, "ms": ms, "points_per_s": len(pts) / ms * 1e3,
                     "kernel_frac_of_its_roofline": rf["frac"], "parity_ok": ok}
        del km, D, I
    return out


def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    import proqa_b200 as pq
    from proqa_b200.sharded import ShardedIndexFlat, auto_row_shards

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; proqa_b200 has no CPU path (use --impl reference for the host baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    nq, N, k = wl["nq"], wl["rows"], wl["k"]
    stream = torch.cuda.current_stream()
    xq_np = host_queries(nq)                                   # pageable, as eval_retrieval.py:99 hands it over
    xq_host = torch.from_numpy(xq_np).pin_memory()
    q = xq_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(t):
        if world > 1:
            tt = torch.tensor([t], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())
        return t

    # layouts to measure: the requested one; with several GPUs and no explicit request, both pure row sharding (north_star (4))
    # and the fewest-row-shards layout — the faster one is the headline, both are reported
    if world == 1:
        layouts = [1]
    elif args.row_shards is None:
        layouts = sorted({world, auto_row_shards(world, N)}, reverse=True)
    else:
        layouts = [auto_row_shards(world, N) if args.row_shards == "auto" else (world if args.row_shards == "rows" else int(args.row_shards))]

    results = {}
    for R in layouts:
        sh = ShardedIndexFlat(128, pq.METRIC_INNER_PRODUCT, device=local_rank, row_shards=R)
        lo, hi = sh.row_bounds(N)
        qlo, qhi = sh.query_bounds(nq)
        ix = sh.local
        if args.tier:
            ix.set_tier(args.tier)
        t_build = time.perf_counter()
        build_shard(ix, lo, hi, dev, sharded=sh, n_global=N)
        sh._first_add, sh.ntotal = False, N
        torch.cuda.synchronize()
        t_build = time.perf_counter() - t_build
        ix.set_stream(stream.cuda_stream)
        D_loc = torch.empty((nq, k), dtype=torch.float32, device=dev)
        I_loc = torch.empty((nq, k), dtype=torch.int64, device=dev)
        D_all = torch.empty((world, nq, k), dtype=torch.float32, device=dev) if world > 1 else None
        I_all = torch.empty((world, nq, k), dtype=torch.int64, device=dev) if world > 1 else None
        D_out, I_out = torch.empty_like(D_loc), torch.empty_like(I_loc)
        D_host = torch.empty((nq, k), dtype=torch.float32).pin_memory()
        I_host = torch.empty((nq, k), dtype=torch.int64).pin_memory()

        def step_device():
            if world == 1:   # results land where the caller asked: no extra copy inside the timed region
                ix.search_device(q.data_ptr(), nq, k, D_out.data_ptr(), I_out.data_ptr())
            else:
                sh.search_device(q, k, D_loc, I_loc, D_all, I_all, D_out, I_out)
            return ix.last_stats[5] + ((5 if sh.R > 1 else 3) if (world > 1 and sh._xchg_on) else (1 if (world > 1 and sh.R > 1) else 0))

        def step_e2e_numpy():
            """The call the reference makes: index.search(pageable float32 array, k) -> new numpy (D, I)  (eval_retrieval.py:104)."""
            return ix.search(xq_np, k) if world == 1 else sh.search(xq_np, k)

        def step_e2e_pinned():
            if world == 1:
                rc = _libmod().lib().pq_index_search(ix._h, nq, ctypes.c_void_p(xq_host.data_ptr()), k, ctypes.c_void_p(D_host.data_ptr()),
                                                     ctypes.c_void_p(I_host.data_ptr()))
                _libmod().check(rc, "search")
            else:
                if rank == 0:
                    q.copy_(xq_host, non_blocking=True)
                dist.broadcast(q, 0)
                sh.search_device(q, k, D_loc, I_loc, D_all, I_all, D_out, I_out)
                if rank == 0:
                    D_host.copy_(D_out, non_blocking=True)
                    I_host.copy_(I_out, non_blocking=True)
                torch.cuda.synchronize()

        # ---- parity gate: >= 256 queries against the fp64 ground truth (checker: torch float64) ---------
        step_device()
        torch.cuda.synchronize()
        nchk = min(nq, 256 if k <= 1024 else 32)
        contributes = sh.rq == 0                      # the ranks of query group 0 together hold every row exactly once
        Dt, It, own = truth_topk_fp64(q[:nchk], lo, hi, N, k, dev, ids=I_out[:nchk])
        if world > 1:
            if not contributes:
                own.zero_()
            dist.all_reduce(own)
            Dt_all = torch.empty((world, nchk, k), dtype=torch.float64, device=dev)
            It_all = torch.empty((world, nchk, k), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(Dt_all.view(world * nchk, k), Dt.contiguous())
            dist.all_gather_into_tensor(It_all.view(world * nchk, k), It.contiguous())
            cat_d = Dt_all[:sh.R].permute(1, 0, 2).reshape(nchk, -1)
            cat_i = It_all[:sh.R].permute(1, 0, 2).reshape(nchk, -1)
            top = torch.topk(cat_d, k, dim=1)
            Dt, It = top.values, torch.gather(cat_i, 1, top.indices)
        parity_ok, parity_info = parity_gate(D_out[:nchk], I_out[:nchk], Dt, It, own)

        # ---- timed region: device-resident -------------------------------------------------------------
        for _ in range(args.warmup):
            step_device()
        sampler = ClockSampler(local_rank)
        barrier()
        if rank == 0:
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches, rerun_q, exchanges = 0, 0, 0
        barrier()
        ev0.record(stream)
        for _ in range(args.steps):
            launches += step_device()
            rerun_q += ix.last_stats[1]
            exchanges += ix.last_stats[9]
        ev1.record(stream)
        barrier()
        t_dev = max_over_ranks(ev0.elapsed_time(ev1) / 1e3)
        clocks = sampler.stop() if rank == 0 else None

        # ---- end to end: host buffers in, host buffers out ----------------------------------------------
        e2e = {}
        for name, fn in (("numpy_pageable", step_e2e_numpy), ("pinned", step_e2e_pinned)):
            for _ in range(max(1, args.warmup // 2)):
                fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                fn()
            barrier()
            e2e[name] = max_over_ranks(time.perf_counter() - t0)

        # ---- phase breakdown of one multi-GPU step (CUDA events on the launching stream; diagnostic) ----
        phases = None
        if world > 1:
            n_loc = qhi - qlo
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            acc = [0.0, 0.0, 0.0, 0.0]
            reps = 3
            for _ in range(reps):
                barrier()
                sh._share_begin(n_loc)
                evs[0].record(stream)
                ix.search_device(q[qlo:qhi].data_ptr(), n_loc, k, D_loc.data_ptr(), I_loc.data_ptr())
                evs[1].record(stream)
                if sh._xchg_on:
                    sh._exchange(D_loc[:n_loc], I_loc[:n_loc], nq, k, D_out, I_out)
                    evs[2].record(stream)
                    evs[3].record(stream)
                else:
                    if sh.R > 1:
                        Da = D_all.view(-1)[: sh.R * n_loc * k].view(sh.R * n_loc, k)
                        Ia = I_all.view(-1)[: sh.R * n_loc * k].view(sh.R * n_loc, k)
                        dist.all_gather_into_tensor(Da, D_loc[:n_loc], group=sh.row_group)
                        dist.all_gather_into_tensor(Ia, I_loc[:n_loc], group=sh.row_group)
                    evs[2].record(stream)
                    if sh.R > 1:
                        sh._merge_device(Da, Ia, n_loc, k, D_out[:n_loc], I_out[:n_loc])
                    evs[3].record(stream)
                    if sh.Q > 1:
                        sh._gather_slices(D_loc[:n_loc], I_loc[:n_loc], nq, k)
                evs[4].record(stream)
                torch.cuda.synchronize()
                for j in range(4):
                    acc[j] += evs[j].elapsed_time(evs[j + 1]) / reps
            if sh._xchg_on:
                phases = {"local_search_ms": max_over_ranks(acc[0]), "peer_memory_exchange_and_merge_ms": max_over_ranks(acc[1]),
                          "exchange": "pq_xchg: scatter to the merging rank, merge of 1/R of the queries, result stored into every rank's HBM (NVLink P2P stores)"}
            else:
                phases = {"local_search_ms": max_over_ranks(acc[0]), "row_group_all_gather_ms": max_over_ranks(acc[1]),
                          "merge_kernel_ms": max_over_ranks(acc[2]), "query_group_all_gather_ms": max_over_ranks(acc[3]),
                          "exchange": "NCCL all-gather + merge kernel on every rank (PROQA_B200_XCHG=0)"}
            phases.update({"local_search_launches": int(ix.last_stats[5]), "threshold_exchanges_in_time_per_search": int(ix.last_stats[9]),
                           "note": "max over ranks of CUDA-event times on the launching stream, one step taken apart"})

        roofline = device_roofline(ix, step_device, 2.0 * (qhi - qlo) * (hi - lo) * 128, 512.0 * (hi - lo), t_dev,
                                   traffic_name=("c2" if (args.workload == "c2" and world == 1) else None))
        results[R] = dict(R=R, Q=world // R, t_dev=t_dev, launches=launches, rerun_q=rerun_q, exchanges=exchanges, e2e=e2e, phases=phases,
                          roofline=roofline, parity_ok=parity_ok, parity_info=parity_info, nchk=nchk, clocks=clocks, t_build=t_build,
                          tensor_path=ix.last_stats[3] > 0)
        if R != layouts[-1]:
            sh.close()
            del sh, ix, D_loc, I_loc, D_all, I_all, D_out, I_out
            torch.cuda.empty_cache()

    sh.close()
    best = min(results.values(), key=lambda r: r["t_dev"] if r["parity_ok"] else float("inf"))
    sweep = None
    if world == 1 and args.workload == "c2" and not args.no_sweep and args.rows is None:
        sweep = sweep_lines(args, ix, dev, stream, N, local_rank)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- index load path: index.add() of a pageable host array (eval_retrieval.py:100,103), GB/s -------------------
    add_info = None
    if world == 1:
        rows_add = 2_000_000
        xh = host_corpus_sample(rows_add)
        scratch = pq.IndexFlatIP(128, local_rank)
        scratch.add(xh[:1000])          # device init, allocations
        scratch.reset()
        t_best = None
        for _ in range(2):
            scratch.reset()
            t0 = time.perf_counter()
            scratch.add(xh)
            dt = time.perf_counter() - t0
            t_best = dt if t_best is None else min(t_best, dt)
        add_info = {"rows": rows_add, "seconds": t_best, "host_gbs": rows_add * 512 / t_best / 1e9,
                    "note": "pageable numpy array -> pinned double-buffered H2D + norms/bf16 preparation, per index.add call"}
        del scratch, xh
    cpu = cpu_baseline(wl) if (world == 1 and not args.no_cpu_baseline) else None
    t_dev, parity_ok = best["t_dev"], best["parity_ok"]
    ms = t_dev / args.steps * 1e3
    cfg = workload_config(args, wl)
    if world > 1:
        cfg["parallelism"] = f"{best['R']} row shards x {best['Q']} query groups"
    t_e2e = best["e2e"]["numpy_pageable"]
    out = {
        "metric": "queries_per_sec", "value": (nq * args.steps / t_dev) if parity_ok else None, "unit": "queries/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16 filter + f32 rescoring" if best["tensor_path"] else "f32",
        "data": "synthetic", "config": cfg,
        "corpus_gbs": N * 512 * args.steps / t_dev / 1e9,
        "e2e": {"value": nq * args.steps / t_e2e, "unit": "queries/s", "h2d_bytes_per_step": nq * 512, "d2h_bytes_per_step": nq * k * 12,
                "ms_per_step": t_e2e / args.steps * 1e3,
                "call": "IndexFlat.search(pageable numpy float32 [nq,128], k) -> new numpy (D, I), as eval_retrieval.py:104 calls it"
                        if world == 1 else "ShardedIndexFlat.search(pageable numpy queries, k) on every rank -> numpy (D, I)",
                "pinned_buffers_ms_per_step": best["e2e"]["pinned"] / args.steps * 1e3},
        "gpu_launches": int(best["launches"]),
        "roofline": best["roofline"],
        "cpu_baseline": cpu,
        "clocks": best["clocks"],
        "parity": dict(best["parity_info"], queries_checked=best["nchk"], ok=parity_ok, against="fp64 brute force (torch, checker only)",
                       fp32_rerun_queries_per_step=best["rerun_q"] / max(1, args.steps)),
        "index_build_s": best["t_build"],
        "index_add": add_info,
        "multi_gpu_phases": best["phases"],
    }
    if world > 1:
        out["layouts"] = {("rows" if r["R"] == world else f"R{r['R']}xQ{r['Q']}"): {
            "row_shards": r["R"], "query_groups": r["Q"], "ms_per_step": r["t_dev"] / args.steps * 1e3,
            "value": nq * args.steps / r["t_dev"] if r["parity_ok"] else None, "parity_ok": r["parity_ok"],
            "e2e_ms_per_step": r["e2e"]["numpy_pageable"] / args.steps * 1e3, "phases": r["phases"],
            "kernel_frac": r["roofline"]["frac"], "filter_kernel_ms": r["roofline"]["kernel_ms_per_step"],
            "other_kernels_ms": r["roofline"].get("other_kernels_ms_per_step"), "gpu_launches": int(r["launches"]),
            "threshold_exchanges_in_time": int(r["exchanges"])} for r in results.values()}
        out["layouts"]["headline"] = "rows" if best["R"] == world else f"R{best['R']}xQ{best['Q']}"
    if sweep is not None:
        out["sweep"] = sweep
    if not parity_ok:
        out["error"] = "parity gate failed: no speed reported"
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _libmod():
    from proqa_b200 import _lib
    return _lib


def run_kmeans_assign(args, wl):
    """C4: the k = 1 assignment search of group_paras.py:45,51 — many points (query side) against few centroids (index side).
    Points are sharded over ranks (data parallel, no exchange); a step is one assignment pass over all points."""
    import torch
    import torch.distributed as dist
    import proqa_b200 as pq
    from proqa_b200.sharded import shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; proqa_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_points, n_cent = wl["nq"], wl["rows"]
    metric = pq.METRIC_L2 if args.metric == "l2" else pq.METRIC_INNER_PRODUCT
    lo, hi = shard_bounds(n_points, world, rank)
    g = torch.Generator(device=dev)
    g.manual_seed(777)
    cents = torch.randn((n_cent, 128), generator=g, device=dev, dtype=torch.float32)
    pts = torch.empty((hi - lo, 128), device=dev, dtype=torch.float32)
    for c in range(lo // CHUNK, (hi + CHUNK - 1) // CHUNK):  # same counter-based stream whatever the rank count
        g.manual_seed(4321 + c)
        rows = min(CHUNK, n_points - c * CHUNK)
        x = torch.randn((rows, 128), generator=g, device=dev, dtype=torch.float32)
        a, b = max(lo, c * CHUNK) - c * CHUNK, min(hi, c * CHUNK + rows) - c * CHUNK
        pts[max(lo, c * CHUNK) - lo: max(lo, c * CHUNK) - lo + (b - a)] = x[a:b]
        del x
    ix = pq.IndexFlat(128, metric, local_rank)
    if args.tier:
        ix.set_tier(args.tier)
    torch.cuda.synchronize()
    ix.add_device(cents.data_ptr(), n_cent)
    stream = torch.cuda.current_stream()
    ix.set_stream(stream.cuda_stream)
    n_loc = hi - lo
    D = torch.empty((n_loc, 1), dtype=torch.float32, device=dev)
    I = torch.empty((n_loc, 1), dtype=torch.int64, device=dev)

    def step():
        ix.search_device(pts.data_ptr(), n_loc, 1, D.data_ptr(), I.data_ptr())
        return ix.last_stats[5]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step()
    torch.cuda.synchronize()
    nchk = min(n_loc, 4096)
    p64, c64 = pts[:nchk].double(), cents.double()
    S = p64 @ c64.T
    if metric == pq.METRIC_L2:
        dist2 = (p64 * p64).sum(1, keepdim=True) + (c64 * c64).sum(1)[None, :] - 2.0 * S
        best, arg = dist2.min(1)
    else:
        best, arg = S.max(1)
    got = D[:nchk, 0].double()
    scale = torch.maximum(best.abs(), torch.full_like(best, 1e-3))
    score_ok = bool(((got - best).abs() <= 1e-4 * scale + 1e-6).all())
    frac_equal = float((I[:nchk, 0] == arg).double().mean())
    parity_ok = score_ok and frac_equal > 0.995

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        launches += step()
    ev1.record(stream)
    barrier()
    t_dev = ev0.elapsed_time(ev1) / 1e3
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tt = torch.tensor([t_dev], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev = float(tt.item())

    # end to end: every step copies a bounded slice of points from pinned host memory and reads the assignment back
    e2e_n = min(n_loc, 2_000_000)
    pts_host = pts[:e2e_n].cpu().pin_memory()
    I_host = torch.empty((e2e_n, 1), dtype=torch.int64).pin_memory()
    D_host = torch.empty((e2e_n, 1), dtype=torch.float32).pin_memory()
    from proqa_b200 import _lib

    def step_e2e():
        rc = _lib.lib().pq_index_search(ix._h, e2e_n, ctypes.c_void_p(pts_host.data_ptr()), 1, ctypes.c_void_p(D_host.data_ptr()),
                                       ctypes.c_void_p(I_host.data_ptr()))
        _lib.check(rc, "search")
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps // 2)):
        step_e2e()
    barrier()
    t_e2e = (time.perf_counter() - t0) / max(1, args.steps // 2)

    ix.set_profile(True)
    kern_us = []
    for _ in range(2):
        step()
        kern_us.append(ix.last_stats[7])
    ix.set_profile(False)
    st = ix.last_stats
    kernel_s = float(np.mean(kern_us)) * 1e-6
    peaks = load_peaks()
    flops = 2.0 * n_loc * n_cent * 128
    long_run = t_dev > 1.0
    peak = peaks["bf16_tflops_sustained"] if long_run else peaks["bf16_tflops"]
    achieved = flops / kernel_s / 1e12 if kernel_s > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": "pq_mma_filter_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": None, "peak_source": f"{peaks['source']} cuBLAS bf16 ({'sustained' if long_run else 'burst'})",
                "algorithmic": "256 flop per (point,centroid) score", "launches_per_step": int(st[3]), "kernel_ms_per_step": kernel_s * 1e3}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle
        ns = 200_000
        xs = pts[:ns].cpu().numpy()
        fo = oracle.FaissFlatOracle(128, 1 if metric == pq.METRIC_L2 else 0)
        fo.add(cents.cpu().numpy())
        t0 = time.perf_counter()
        fo.search(xs, 1)
        dt = time.perf_counter() - t0
        cpu = {"value": ns / dt, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"first {ns} of {n_points} points x all {n_cent} centroids in {dt:.2f}s; FAISS-1.6.3 restatement (OpenBLAS sgemm + heap), not FAISS"}
    out = {"metric": "queries_per_sec", "value": (n_points * args.steps / t_dev) if parity_ok else None, "unit": "queries/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "bf16 filter + f32 rescoring", "data": "synthetic",
           "config": {"workload": wl["desc"], "name": args.workload, "points": n_points, "centroids": n_cent, "d": 128, "k": 1,
                      "metric": args.metric, "l2_flush": "points streamed per step (%.1f GB) exceed L2" % (n_loc * 512 / 1e9)},
           "corpus_gbs": n_points * 512 * args.steps / t_dev / 1e9,
           "e2e": {"value": e2e_n * world / t_e2e, "unit": "queries/s", "h2d_bytes_per_step": e2e_n * 512, "d2h_bytes_per_step": e2e_n * 12,
                   "sample": f"{e2e_n} points per step per rank through pq_index_search (pinned host in/out)"},
           "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
           "parity": {"points_checked": nchk, "ok": parity_ok, "ids_equal_frac": frac_equal, "against": "fp64 brute force (torch, checker only)",
                      "fp32_rerun_queries_per_step": int(st[1])}}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def load_traffic(kernel):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=None)
    ap.add_argument("--nq", type=int, default=None)
    ap.add_argument("--k", type=int, default=None)
    ap.add_argument("--tier", default=None, choices=[None, "auto", "fp32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the S0 / C1 / C4 lines the default (c2, 1 GPU) run adds")
    ap.add_argument("--row-shards", default=None,
                    help="multi-GPU layout: number of corpus row shards R (world = R x query groups); 'rows' = one shard per GPU (pure row "
                         "sharding + NCCL merge, north_star (4)), 'auto' = fewest shards whose slice of the corpus fits the per-GPU budget; "
                         "default: measure both, headline the faster, report both under \"layouts\"")
    ap.add_argument("--metric", default="l2", choices=["ip", "l2"], help="c4 only (group_paras.py default is L2; --spherical is IP)")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    for key in ("rows", "nq", "k"):
        if getattr(args, key) is not None:
            wl[key] = getattr(args, key)
            wl["desc"] += f" [{key}={wl[key]} override]"
    if args.impl == "reference":
        run_reference(args, wl)
    elif wl.get("kind") == "kmeans":
        run_kmeans_assign(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
