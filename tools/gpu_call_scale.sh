#!/bin/bash
# ONE 8-GPU call (charged 8x): C2 at N=8 / 4 / 2 (default exchange) + the round-1 exchange at N=8, C4 on 4 GPUs, C5 on 8 GPUs.
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_call_scale.sh'
cd "$(dirname "$0")/.."
OUT=gpurun_out/scale
mkdir -p "$OUT"
run() {   # name, gpus, extra flags
  local name=$1 n=$2; shift 2
  local t0=$(date +%s)
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29513 \
      bench.py --gpus "$n" --steps 30 --warmup 5 --no-cpu-baseline --no-profile "$@" > "$OUT/$name.json" 2> "$OUT/$name.err"
  local rc=$?
  python - "$OUT/$name.json" "$name" "$rc" "$(( $(date +%s) - t0 ))" <<'PY' | tee -a "$OUT/summary.log"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"{sys.argv[2]}: {d['ms_per_step']:.3f} ms/step {d['value']:.1f} samples/s n_gpus {d['n_gpus']} loss {d['loss']} (rc {sys.argv[3]}, {sys.argv[4]} s wall)")
except Exception as e:
    print(f"{sys.argv[2]}: FAILED rc {sys.argv[3]} ({e})")
PY
}
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > "$OUT/smi.txt" 2>&1
: > "$OUT/summary.log"
run c2_n8 8
run c2_n8_round1_exchange 8 --whole-arena-reduce --dense-embedding-reduce
run c2_n4 4
run c2_n2 2
run c4_n4 4 --workload c4 --steps 10 --warmup 3
run c5_n8 8 --workload c5 --steps 10 --warmup 3
echo "=== done" | tee -a "$OUT/summary.log"
