#!/bin/bash
# 1-GPU call: parallel split-K weight-gradient GEMMs (TMA reduce-add) on hardware: tests, C2 bench + A/B with the switch off, C3.
cd "$(dirname "$0")/.."
OUT=gpurun_out/splitk
mkdir -p "$OUT"
line() {
python - "$1" <<'PY' | tee -a "$OUT/summary.log"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(f"  {d['ms_per_step']:.3f} ms/step {d['value']:.1f} samples/s e2e {d['e2e']['value']:.1f} gemm frac {r.get('frac')} lib ms {r.get('library_kernel_ms_per_step')} "
          f"profiled step {r.get('step_ms_under_profiler_events')} xattn frac {(d.get('xattn') or {}).get('frac')}")
    for i in (r.get("instantiations") or []):
        print("   ", i["tag"], i["launches_per_step"], round(i["avg_launch_ms"] * 1e3, 1), "us", round(i["achieved"]), "TF", round(i["frac"], 3))
    for k, v in list((d.get("kernels") or {}).items())[:24]:
        print("    k", k, v["launches"], round(v["ms_per_step"], 3))
except Exception as e:
    print("  FAILED:", e)
PY
}
echo "=== tests" | tee "$OUT/summary.log"
for f in tests/test_gpu_gemm.py tests/test_gpu_modules.py tests/test_gpu_model.py tests/test_gpu_training.py; do
  timeout 900 python -m pytest "$f" -q -m gpu --tb=short 2>&1 | tail -6 | tee -a "$OUT/summary.log"
done
echo "=== bench c2" | tee -a "$OUT/summary.log"
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > "$OUT/bench_c2.json" 2> "$OUT/bench_c2.err"; line "$OUT/bench_c2.json"
echo "=== bench c2 dw_splitk=0" | tee -a "$OUT/summary.log"
FM_B200_OPTS="dw_splitk=0" timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-profile > "$OUT/bench_c2_nosplit.json" 2> "$OUT/bench_c2_nosplit.err"; line "$OUT/bench_c2_nosplit.json" | head -1
echo "=== bench c3" | tee -a "$OUT/summary.log"
timeout 600 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline > "$OUT/bench_c3.json" 2> "$OUT/bench_c3.err"; line "$OUT/bench_c3.json" | head -8
timeout 200 python tools/gemm_trace.py dw1 > "$OUT/gemm_trace_dw.txt" 2>&1; grep "^cta  0\|^==" "$OUT/gemm_trace_dw.txt" | cut -c1-300 | tee -a "$OUT/summary.log"
echo "=== done" | tee -a "$OUT/summary.log"
