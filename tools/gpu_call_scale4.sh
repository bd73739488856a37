#!/bin/bash
# 8-GPU, ONE run: C2 at N=8 with NCCL_PROTO=Simple (the CUPTI timeline showed 45 ring all-reduces in the LL protocol, ~410 us each,
# active for 9.3 ms of an 11.1 ms step: profiles/r02_call13_scale3/dp_timeline_n8.json).
cd "$(dirname "$0")/.."
OUT=gpurun_out/scale4
mkdir -p "$OUT"
NCCL_PROTO=Simple timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu-baseline --no-profile > "$OUT/c2_n8_proto_simple.json" 2> "$OUT/c2_n8_proto_simple.err"
python - "$OUT/c2_n8_proto_simple.json" <<'PY' | tee "$OUT/summary.log"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"c2_n8 NCCL_PROTO=Simple: {d['ms_per_step']:.3f} ms/step {d['value']:.1f} samples/s n_gpus {d['n_gpus']} loss {d['loss']}")
except Exception as e:
    print(f"FAILED ({e})")
PY
tail -3 "$OUT/c2_n8_proto_simple.err" | cut -c1-300 >> "$OUT/summary.log"
