#!/bin/bash
# 2-GPU (or N-GPU) call: LayerNorm kernels on hardware first, then the data-parallel A/B (tools/validate_dp.sh).
cd "$(dirname "$0")/.."
OUT=gpurun_out/dp
mkdir -p "$OUT"
N=${N:-2}
echo "=== LN + module tests on hardware" | tee "$OUT/pre_n${N}.log"
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_modules.py tests/test_gpu_model.py tests/test_gpu_training.py -q -m gpu --tb=short 2>&1 | tail -8 | tee -a "$OUT/pre_n${N}.log"
echo "=== 1-GPU bench (new LayerNorm kernels)" | tee -a "$OUT/pre_n${N}.log"
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"
python - "$OUT/bench_n1.json" <<'PY' | tee -a "$OUT/pre_n${N}.log"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(f"  {d['ms_per_step']:.3f} ms/step {d['value']:.1f} samples/s gemm frac {r.get('frac')} lib ms {r.get('library_kernel_ms_per_step')}")
    for k, v in (d.get("kernels") or {}).items():
        if k.startswith("ln_"):
            print("    k", k, v["launches"], round(v["ms_per_step"], 3))
except Exception as e:
    print("  FAILED:", e)
PY
timeout 300 python tools/decode_bench.py 2>&1 | tail -1 | tee "$OUT/decode_bench.json" | cut -c1-400
N=$N bash tools/validate_dp.sh
