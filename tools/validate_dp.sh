#!/bin/bash
# Data-parallel A/B on 2 GPUs (charged 2x):  /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/validate_dp.sh'
# default exchange vs --split-embedding (parallel.SplitEmbeddingGrad) on the validated build, then the staging build without
# and with (--split-embedding --per-layer-reduce); the losses must agree, the step time should drop.
cd "$(dirname "$0")/.."
OUT=gpurun_out/dp
mkdir -p "$OUT"
N=${N:-2}
for mode in default split next next_all next_all_reserve16 next_all_reserve32; do
  flag=""; variant=""; opts=""
  [ "$mode" = next_all_reserve16 ] && { variant=next; flag="--split-embedding --per-layer-reduce"; opts="sm_reserve=16"; }
  [ "$mode" = next_all_reserve32 ] && { variant=next; flag="--split-embedding --per-layer-reduce"; opts="sm_reserve=32"; }
  [ "$mode" = split ] && flag="--split-embedding"
  [ "$mode" = next ] && variant=next
  [ "$mode" = next_all ] && { variant=next; flag="--split-embedding --per-layer-reduce"; }
  FM_B200_VARIANT=$variant FM_B200_OPTS=$opts timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus "$N" --steps 30 --warmup 5 --no-profile $flag > "$OUT/bench_n${N}_${mode}.json" 2> "$OUT/bench_n${N}_${mode}.err"
  python - "$OUT/bench_n${N}_${mode}.json" "$mode" <<'PY' || head -c 2000 "$OUT/bench_n${N}_${mode}.err"
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], d["ms_per_step"], "ms/step", d["value"], "samples/s  loss", d["loss"])
PY
done
