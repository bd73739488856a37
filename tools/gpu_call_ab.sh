#!/bin/bash
# 1-GPU call: C2 bench with the single-stream CUPTI profile, A/B lines (side stream off, PDL off, GEMM grouping off), refreshed in-situ
# ncu captures of the kernels that carry the step.
cd "$(dirname "$0")/.."
OUT=gpurun_out/ab
mkdir -p "$OUT"
line() {
python - "$1" <<'PY' | tee -a "$OUT/summary.log"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(f"  {d['ms_per_step']:.3f} ms/step {d['value']:.1f} samples/s e2e {d['e2e']['value']:.1f} gemm frac {r.get('frac')} lib ms {r.get('library_kernel_ms_per_step')} "
          f"profiled step {r.get('step_ms_under_profiler_events')} xattn frac {(d.get('xattn') or {}).get('frac')} traffic {r.get('traffic')}")
    print("   cfg:", {k: v for k, v in d["config"].items() if "cupti" in k or "error" in k or "options" in k})
    for i in (r.get("instantiations") or []):
        print("   ", i["tag"], i["launches_per_step"], round(i["avg_launch_ms"] * 1e3, 1), "us", round(i["achieved"]), "TF", round(i["frac"], 3))
    for k, v in list((d.get("kernels") or {}).items())[:36]:
        print("    k", k, v["launches"], round(v["ms_per_step"], 3))
except Exception as e:
    print("  FAILED:", e)
PY
}
echo "=== bench c2 (defaults)" | tee "$OUT/summary.log"
timeout 600 python bench.py --steps 30 --warmup 5 > "$OUT/bench_c2.json" 2> "$OUT/bench_c2.err"; line "$OUT/bench_c2.json"
for ab in "side_stream=0" "pdl=0" "gemm_group=0" "side_stream=0,pdl=0"; do
  echo "=== bench c2 FM_B200_OPTS=$ab" | tee -a "$OUT/summary.log"
  FM_B200_OPTS="$ab" timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-profile > "$OUT/bench_c2_${ab//[=,]/_}.json" 2> "$OUT/bench_c2_${ab//[=,]/_}.err"
  line "$OUT/bench_c2_${ab//[=,]/_}.json" | head -1
done
echo "=== ncu in situ" | tee -a "$OUT/summary.log"
bash tools/ncu_top_kernels.sh 2>&1 | tail -8 | tee -a "$OUT/summary.log"
mkdir -p "$OUT/ncu"
for f in gpurun_out/ncu/*.ncu-rep; do
  n=$(basename "$f" .ncu-rep)
  ncu -i "$f" --page raw --csv 2>/dev/null | python tools/ncu_pick.py > "$OUT/ncu/$n.txt" 2>&1
done
rm -f gpurun_out/ncu/*.ncu-rep       # keep the text extracts only (the reports are 8 MB each)
echo "=== done" | tee -a "$OUT/summary.log"
