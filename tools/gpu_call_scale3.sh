#!/bin/bash
# Last 8-GPU call (charged 8x, keep it short): CUPTI timeline of the data-parallel step on rank 0 (what is exposed, how long NCCL
# kernels sit on SMs), and the bucketed exchange A/B.
cd "$(dirname "$0")/.."
OUT=gpurun_out/scale3
mkdir -p "$OUT"
: > "$OUT/summary.log"
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 tools/dp_timeline.py 2> "$OUT/dp_timeline.err" | tail -1 | tee "$OUT/dp_timeline_n8.json" | cut -c1-900 | tee -a "$OUT/summary.log"
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
run c2_n8_bucket3 8 --reduce-bucket 3
run c2_n8_bucket6 8 --reduce-bucket 6
echo "=== done" | tee -a "$OUT/summary.log"
