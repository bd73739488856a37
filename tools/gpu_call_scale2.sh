#!/bin/bash
# 8-GPU call (charged 8x, keep it short): C2 at N=8 and N=4 with fp32 and bf16 gradient wire formats, final kernels.
cd "$(dirname "$0")/.."
OUT=gpurun_out/scale2
mkdir -p "$OUT"
: > "$OUT/summary.log"
run() {
  local name=$1 n=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus "$n" --steps 30 --warmup 5 --no-cpu-baseline --no-profile "$@" > "$OUT/$name.json" 2> "$OUT/$name.err"
  python - "$OUT/$name.json" "$name" <<'PY' | tee -a "$OUT/summary.log"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"{sys.argv[2]}: {d['ms_per_step']:.3f} ms/step {d['value']:.1f} samples/s n_gpus {d['n_gpus']} loss {d['loss']}")
except Exception as e:
    print(f"{sys.argv[2]}: FAILED ({e})")
PY
}
run c2_n8 8
run c2_n8_bf16wire 8 --bf16-wire
run c2_n4 4
run c2_n4_bf16wire 4 --bf16-wire
echo "=== done" | tee -a "$OUT/summary.log"
