#!/bin/bash
# 1-GPU call validating the TMA-based GEMM epilogue (inputs by TMA load one item ahead, outputs by TMA store):
# bring-up probe, GPU test files one process each, bench C2 (CUPTI per-kernel durations), per-CTA timelines, C3/C4 lines.
cd "$(dirname "$0")/.."
OUT=gpurun_out/epi
mkdir -p "$OUT"
echo "=== gemm_diag" | tee "$OUT/summary.log"
timeout 300 python tools/gemm_diag.py 2>&1 | tail -4 | tee -a "$OUT/summary.log"
for f in tests/test_gpu_gemm.py tests/test_gpu_ops.py tests/test_gpu_modules.py tests/test_gpu_model.py tests/test_gpu_training.py; do
  echo "=== $f" | tee -a "$OUT/summary.log"
  timeout 900 python -m pytest "$f" -q -m gpu --tb=short 2>&1 | tail -12 | tee -a "$OUT/summary.log"
done
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | tee -a "$OUT/summary.log"
for wl in c2 c3 c4; do
  echo "=== bench $wl" | tee -a "$OUT/summary.log"
  extra="--steps 30 --warmup 5"; [ "$wl" != c2 ] && extra="--steps 10 --warmup 3"
  timeout 600 python bench.py --workload $wl $extra --no-cpu-baseline > "$OUT/bench_$wl.json" 2> "$OUT/bench_$wl.err"
  python - "$OUT/bench_$wl.json" <<'PY' | tee -a "$OUT/summary.log"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(f"  {d['ms_per_step']:.3f} ms/step {d['value']:.1f} samples/s e2e {d['e2e']['value']:.1f} gemm frac {r.get('frac')} lib ms {r.get('library_kernel_ms_per_step')} "
          f"profiled step {r.get('step_ms_under_profiler_events')} xattn frac {(d.get('xattn') or {}).get('frac')}")
    print("   timing:", (r.get("timing") or "")[:90], "| cfg:", {k: v for k, v in d["config"].items() if "cupti" in k or "error" in k})
    for i in (r.get("instantiations") or []):
        print("   ", i["tag"], i["launches_per_step"], round(i["avg_launch_ms"] * 1e3, 1), "us", round(i["achieved"]), "TF", round(i["frac"], 3))
    for k, v in list((d.get("kernels") or {}).items())[:34]:
        print("    k", k, v["launches"], round(v["ms_per_step"], 3))
except Exception as e:
    print("  FAILED:", e)
PY
done
echo "=== per-CTA timelines" | tee -a "$OUT/summary.log"
timeout 300 python tools/gemm_trace.py ffw1 ffw2 dact dx > "$OUT/gemm_trace.txt" 2>&1
grep "^cta  0\|^==" "$OUT/gemm_trace.txt" | cut -c1-330 | tee -a "$OUT/summary.log"
echo "=== done" | tee -a "$OUT/summary.log"
