#!/bin/bash
# Closing 1-GPU call of round 2 (a few minutes): the -m gpu suite and smoke() on the final tree, the C2 line with the shipped
# defaults, and the ncu launch list of the same bench command (the recipe's `--metrics gpu__time_duration.sum --clock-control
# none` pass) summarised per kernel for one step.
cd "$(dirname "$0")/.."
OUT=gpurun_out/close
mkdir -p "$OUT"
: > "$OUT/summary.log"
echo "=== pytest -m gpu" | tee -a "$OUT/summary.log"
timeout 170 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee -a "$OUT/summary.log"
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a "$OUT/summary.log"
echo "=== bench c2 (defaults)" | tee -a "$OUT/summary.log"
timeout 150 python bench.py --steps 30 --warmup 5 > "$OUT/bench_c2.json" 2> "$OUT/bench_c2.err"
python - "$OUT/bench_c2.json" <<'PY' | tee -a "$OUT/summary.log"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(f"c2: {d['ms_per_step']:.3f} ms/step {d['value']:.1f} samples/s e2e {d['e2e']['value']:.1f} gemm frac {r.get('frac')} "
          f"xattn frac {(d.get('xattn') or {}).get('frac')} launches/step {d['gpu_launches'] / d['steps']:.0f} clocks {d.get('clocks')} cpu {d.get('cpu_baseline')}")
except Exception as e:
    print("c2 FAILED:", e)
PY
echo "=== ncu launch list (bench.py --steps 2 --warmup 3, eager so that every kernel is its own launch)" | tee -a "$OUT/summary.log"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file "$OUT/launches.csv" \
  python bench.py --steps 2 --warmup 3 --no-graph --no-profile --no-cpu-baseline > "$OUT/bench_under_ncu.log" 2>&1
python tools/summarize_launches.py "$OUT/launches.csv" > "$OUT/ncu_launch_summary.txt" 2>&1
head -12 "$OUT/ncu_launch_summary.txt" | tee -a "$OUT/summary.log"
gzip -f "$OUT/launches.csv"
echo "=== done" | tee -a "$OUT/summary.log"
