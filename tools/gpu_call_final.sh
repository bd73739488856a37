#!/bin/bash
# Final 1-GPU evidence call of round 2: whole GPU suite, smoke, C2 bench (+ A/B of the operand L2 prefetch), C3/C4/C5 lines, parity
# report, decode bench, per-CTA timelines.  Everything lands in gpurun_out/final/.
cd "$(dirname "$0")/.."
OUT=gpurun_out/final
mkdir -p "$OUT"
line() {
python - "$1" "${2:-40}" <<'PY' | tee -a "$OUT/summary.log"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(f"  {d['ms_per_step']:.3f} ms/step {d['value']:.1f} samples/s e2e {d['e2e']['value']:.1f} gemm frac {r.get('frac')} lib ms {r.get('library_kernel_ms_per_step')} "
          f"profiled step {r.get('step_ms_under_profiler_events')} xattn frac {(d.get('xattn') or {}).get('frac')} launches/step {d['gpu_launches'] / d['steps']:.0f} cpu {(d.get('cpu_baseline') or {}).get('value')}")
    n = int(sys.argv[2])
    for i in (r.get("instantiations") or [])[:n]:
        print("   ", i["tag"], i["launches_per_step"], round(i["avg_launch_ms"] * 1e3, 1), "us", round(i["achieved"]), "TF", round(i["frac"], 3))
    for k, v in list((d.get("kernels") or {}).items())[:n]:
        print("    k", k, v["launches"], round(v["ms_per_step"], 3))
except Exception as e:
    print("  FAILED:", e)
PY
}
echo "=== pytest -m gpu" | tee "$OUT/summary.log"
timeout 1200 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -6 | tee -a "$OUT/summary.log"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | tee -a "$OUT/summary.log"
echo "=== bench c2 (defaults)" | tee -a "$OUT/summary.log"
timeout 600 python bench.py --steps 30 --warmup 5 > "$OUT/bench_c2.json" 2> "$OUT/bench_c2.err"; line "$OUT/bench_c2.json"
echo "=== bench c2 FM_B200_OPTS=epi_prefetch=0 (operand L2 prefetch off)" | tee -a "$OUT/summary.log"
FM_B200_OPTS="epi_prefetch=0" timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > "$OUT/bench_c2_noprefetch.json" 2> "$OUT/bench_c2_noprefetch.err"; line "$OUT/bench_c2_noprefetch.json" 8
for wl in c3 c4 c5; do
  echo "=== bench $wl" | tee -a "$OUT/summary.log"
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > "$OUT/bench_$wl.json" 2> "$OUT/bench_$wl.err"; line "$OUT/bench_$wl.json" 7
done
echo "=== reference arm (CPU oracle port, full batch, bounded)" | tee -a "$OUT/summary.log"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference_arm.json" 2> "$OUT/bench_reference_arm.err"; tail -c 600 "$OUT/bench_reference_arm.json" | tee -a "$OUT/summary.log"; echo | tee -a "$OUT/summary.log"
echo "=== parity report" | tee -a "$OUT/summary.log"
timeout 900 python tools/parity_report.py --out "$OUT/r02_parity.md" 2>&1 | tail -22 | tee -a "$OUT/summary.log"
echo "=== decode bench" | tee -a "$OUT/summary.log"
timeout 300 python tools/decode_bench.py 2>&1 | tail -1 | tee "$OUT/decode_bench.json" | cut -c1-500 | tee -a "$OUT/summary.log"
echo "=== per-CTA timelines" | tee -a "$OUT/summary.log"
timeout 300 python tools/gemm_trace.py ffw1 ffw2 dact dx dw1 > "$OUT/gemm_trace.txt" 2>&1
grep "^cta  0\|^==" "$OUT/gemm_trace.txt" | cut -c1-330 | tee -a "$OUT/summary.log"
echo "=== done" | tee -a "$OUT/summary.log"
