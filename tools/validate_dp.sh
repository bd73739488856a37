#!/bin/bash
# Data-parallel A/B on N GPUs (charged Nx):  /usr/local/graft/bin/gpurun --gpus 2 --timeout 1200 -- 'bash tools/validate_dp.sh'
# 1. pytest -m gpu tests/test_gpu_nccl.py  (rank-averaged gradients == single-process gradients on the concatenated batch)
# 2. C2 bench lines: the default exchange (per-layer resampler reduce + split embedding exchange), the round-1 exchange, each
#    of the two switched off, and sm_reserve; the losses must agree.  Results land in gpurun_out/dp/.
cd "$(dirname "$0")/.."
OUT=gpurun_out/dp
mkdir -p "$OUT"
N=${N:-2}
echo "=== NCCL gradient equality" | tee "$OUT/summary_n${N}.log"
timeout 600 python -m pytest tests/test_gpu_nccl.py -q -m gpu --tb=short 2>&1 | tail -8 | tee -a "$OUT/summary_n${N}.log"
for mode in ${MODES:-default bf16wire old whole_arena dense_embedding}; do
  flag=""; opts=""
  [ "$mode" = old ] && flag="--whole-arena-reduce --dense-embedding-reduce"       # round-1 exchange
  [ "$mode" = whole_arena ] && flag="--whole-arena-reduce"
  [ "$mode" = dense_embedding ] && flag="--dense-embedding-reduce"
  [ "$mode" = bf16wire ] && flag="--bf16-wire"
  FM_B200_OPTS=$opts timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus "$N" --steps 30 --warmup 5 --no-profile --no-cpu-baseline $flag > "$OUT/bench_n${N}_${mode}.json" 2> "$OUT/bench_n${N}_${mode}.err"
  python - "$OUT/bench_n${N}_${mode}.json" "$mode" <<'PY' | tee -a "$OUT/summary_n${N}.log" || head -c 2000 "$OUT/bench_n${N}_${mode}.err"
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], d["ms_per_step"], "ms/step", d["value"], "samples/s  loss", d["loss"])
PY
done
