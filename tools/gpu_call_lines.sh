#!/bin/bash
# 1-GPU call: bench lines of every BASELINE configuration with the final defaults (no profile extras, no CPU baseline except C2).
cd "$(dirname "$0")/.."
OUT=gpurun_out/lines
mkdir -p "$OUT"
: > "$OUT/summary.log"
for wl in c2 c3 c4 c5; do
  extra="--steps 10 --warmup 3 --no-cpu-baseline"; [ "$wl" = c2 ] && extra="--steps 30 --warmup 5"
  timeout 600 python bench.py --workload $wl $extra > "$OUT/bench_$wl.json" 2> "$OUT/bench_$wl.err"
  python - "$OUT/bench_$wl.json" $wl <<'PY' | tee -a "$OUT/summary.log"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(f"{sys.argv[2]}: {d['ms_per_step']:.3f} ms/step {d['value']:.1f} samples/s e2e {d['e2e']['value']:.1f} gemm frac {r.get('frac')} lib ms {r.get('library_kernel_ms_per_step')} "
          f"xattn frac {(d.get('xattn') or {}).get('frac')} launches/step {d['gpu_launches'] / d['steps']:.0f} clocks {d.get('clocks')}")
    for i in (r.get("instantiations") or []):
        print("   ", i["tag"], i["launches_per_step"], round(i["avg_launch_ms"] * 1e3, 1), "us", round(i["achieved"]), "TF", round(i["frac"], 3))
except Exception as e:
    print(sys.argv[2], "FAILED:", e)
PY
done
echo "=== done" | tee -a "$OUT/summary.log"
