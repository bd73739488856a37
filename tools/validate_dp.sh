#!/bin/bash
# Data-parallel A/B on N GPUs (charged Nx):  /usr/local/graft/bin/gpurun --gpus 2 --timeout 1200 -- 'bash tools/validate_dp.sh'
# 1. pytest -m gpu tests/test_gpu_nccl.py  (rank-averaged gradients == single-process gradients on the concatenated batch)
# 2. C2 bench lines: default exchange, --split-embedding, --per-layer-reduce, both, both + sm_reserve; the losses must agree,
#    the step time should drop.  Results land in gpurun_out/dp/.
cd "$(dirname "$0")/.."
OUT=gpurun_out/dp
mkdir -p "$OUT"
N=${N:-2}
echo "=== NCCL gradient equality" | tee "$OUT/summary_n${N}.log"
timeout 600 python -m pytest tests/test_gpu_nccl.py -q -m gpu --tb=short 2>&1 | tail -8 | tee -a "$OUT/summary_n${N}.log"
for mode in ${MODES:-default split perlayer all all_reserve8 all_reserve16}; do
  flag=""; opts=""
  [ "$mode" = split ] && flag="--split-embedding"
  [ "$mode" = perlayer ] && flag="--per-layer-reduce"
  [ "$mode" = all ] && flag="--split-embedding --per-layer-reduce"
  [ "$mode" = all_reserve8 ] && { flag="--split-embedding --per-layer-reduce"; opts="sm_reserve=8"; }
  [ "$mode" = all_reserve16 ] && { flag="--split-embedding --per-layer-reduce"; opts="sm_reserve=16"; }
  FM_B200_OPTS=$opts timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus "$N" --steps 30 --warmup 5 --no-profile --no-cpu-baseline $flag > "$OUT/bench_n${N}_${mode}.json" 2> "$OUT/bench_n${N}_${mode}.err"
  python - "$OUT/bench_n${N}_${mode}.json" "$mode" <<'PY' | tee -a "$OUT/summary_n${N}.log" || head -c 2000 "$OUT/bench_n${N}_${mode}.err"
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], d["ms_per_step"], "ms/step", d["value"], "samples/s  loss", d["loss"])
PY
done
