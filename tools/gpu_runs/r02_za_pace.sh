#!/bin/bash
# Producer pacing (pq_mma.cu: pace_*): parity at BASELINE sizes with pacing on, A/B of C2 and the C5 shard (pacing on / off,
# interleaved), DRAM traffic of the last C2 epoch with pacing on (ncu full set).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py tests/test_gpu_largek.py tests/test_gpu_multi.py -m gpu -x -q > $O/za_pytest.log 2>&1
echo "gpu tests exit $?"; tail -3 $O/za_pytest.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms", round(d["ms_per_step"],3), "frac", round(d["roofline"]["frac"],4), "kern ms", round(d["roofline"]["kernel_ms_per_step"],3),
          "parity", d["parity"]["ok"], "e2e ms", round(d["e2e"]["ms_per_step"],3), d.get("clocks"))
except Exception as e:
    print("parse failed", sys.argv[1], e); print(open(sys.argv[1].replace(".json",".err")).read()[-2000:])
PY
}
PROQA_B200_PACE_MIN_TILES=1 PROQA_B200_PACE_SHIFT=3 timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/za_pytest_forced.log 2>&1
echo "gpu parity tests with pacing forced on small epochs: exit $?"; tail -2 $O/za_pytest_forced.log
PROQA_B200_PACE_MIN_TILES=1 PROQA_B200_PACE_SHIFT=2 timeout -s KILL 400 compute-sanitizer --tool memcheck python tools/sanitize_small.py > $O/za_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -3 $O/za_memcheck.log
for rep in 1 2; do
for pace in 1 0; do
PROQA_B200_PACE=$pace timeout -s KILL 300 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline --no-sweep > $O/za_c2_p${pace}_$rep.json 2> $O/za_c2_p${pace}_$rep.err; show $O/za_c2_p${pace}_$rep.json
done
done
for pace in 1 0; do
PROQA_B200_PACE=$pace timeout -s KILL 300 python bench.py --workload c5 --rows 12500000 --steps 5 --warmup 2 --no-cpu-baseline > $O/za_c5_p$pace.json 2> $O/za_c5_p$pace.err; show $O/za_c5_p$pace.json
done
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:pq_mma_filter -s 5 -c 1 -o $O/za_prof_last -f python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline --no-sweep > $O/za_ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu -i $O/za_prof_last.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[2:]:
    for k in ('dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','lts__t_bytes.sum'):
        if k in h: print(k, r[h.index(k)], rows[1][h.index(k)])
"
# epoch schedule sweep (tuning hooks): C2 and C1 at other growth factors / bootstrap sizes
for g in 4 16; do
PROQA_B200_GROWTH=$g timeout -s KILL 300 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline --no-sweep > $O/za_c2_g$g.json 2> $O/za_c2_g$g.err; show $O/za_c2_g$g.json
done
PROQA_B200_BOOT_ROWS=4096 timeout -s KILL 300 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline --no-sweep > $O/za_c2_b4096.json 2> $O/za_c2_b4096.err; show $O/za_c2_b4096.json
for g in 0 16 32 1024; do
PROQA_B200_GROWTH=$g timeout -s KILL 200 python bench.py --workload c1 --steps 50 --warmup 5 --no-cpu-baseline > $O/za_c1_g$g.json 2> $O/za_c1_g$g.err; show $O/za_c1_g$g.json
done
PROQA_B200_BOOT_ROWS=4096 PROQA_B200_GROWTH=16 timeout -s KILL 200 python bench.py --workload c1 --steps 50 --warmup 5 --no-cpu-baseline > $O/za_c1_b4096g16.json 2> $O/za_c1_b4096g16.err; show $O/za_c1_b4096g16.json
