#!/bin/bash
# ONE gpurun call that validates the staging tree (csrc_next/ -> libflamingo_b200_next.so) against the validated build:
#
#   /usr/local/graft/bin/gpurun --timeout 2100 -- 'bash tools/validate_next.sh'     (≈ 25 min on the box)
#
# 1. validated build: pytest -m gpu (baseline sanity) + one bench line
# 2. staging build:   GEMM bring-up probe, every -m gpu test file in its own process (a trapped kernel leaves a sticky
#                     CUDA error: it must not mask the other files), smoke()
# 3. A/B bench lines of the staging build: defaults, then one scheduling switch flipped at a time (FM_B200_OPTS)
# 4. ncu launch list of the staging build's step (kernel shares; never a bench number)
# Everything lands in gpurun_out/next/; every step runs under its own `timeout` so a hang costs minutes, not the box.
cd "$(dirname "$0")/.."
OUT=gpurun_out/next
mkdir -p "$OUT"
STEPS=${STEPS:-30}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/smi.txt" 2>&1

run_bench() {   # name, variant, opts, [extra bench.py flags]
  echo "=== bench $1 (variant='$2' opts='$3' $4)" | tee -a "$OUT/summary.log"
  FM_B200_VARIANT="$2" FM_B200_OPTS="$3" timeout 420 python bench.py --steps "$STEPS" --warmup 5 --no-cpu-baseline $4 \
      > "$OUT/bench_$1.json" 2> "$OUT/bench_$1.err"
  python - "$OUT/bench_$1.json" <<'PY' | tee -a "$OUT/summary.log"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(f"  {d['ms_per_step']:.3f} ms/step  {d['value']:.1f} samples/s  gemm frac {r.get('frac')}  lib ms {r.get('library_kernel_ms_per_step')}  loss {d.get('loss')}")
except Exception as e:
    print("  FAILED:", e)
PY
}

echo "=== [validated] pytest -m gpu" | tee "$OUT/summary.log"
timeout 900 python -m pytest tests -q -m gpu --tb=short -s 2>&1 | grep -v "^$" | tail -60 | tee -a "$OUT/summary.log"
run_bench validated "" ""

export FM_B200_VARIANT=next
echo "=== [next] gemm_diag" | tee -a "$OUT/summary.log"
timeout 300 python tools/gemm_diag.py 2>&1 | tail -25 | tee -a "$OUT/summary.log"
for f in tests/test_gpu_gemm.py tests/test_gpu_ops.py tests/test_gpu_modules.py tests/test_gpu_model.py; do
  echo "=== [next] $f" | tee -a "$OUT/summary.log"
  timeout 900 python -m pytest "$f" -q -m gpu --tb=short 2>&1 | tail -40 | tee -a "$OUT/summary.log"
done
echo "=== [next] smoke" | tee -a "$OUT/summary.log"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee -a "$OUT/summary.log"
echo "=== [next] compute-sanitizer memcheck on smoke() (out-of-bounds global/shared accesses that happen to give right answers)" | tee -a "$OUT/summary.log"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python __graft_entry__.py smoke > "$OUT/memcheck_smoke.log" 2>&1
echo "  exit $? ; $(grep -c 'Invalid\|out of bounds' "$OUT/memcheck_smoke.log") suspicious lines ; $(tail -1 "$OUT/memcheck_smoke.log")" | tee -a "$OUT/summary.log"
unset FM_B200_VARIANT

run_bench next_default next ""
run_bench next_pdl next "pdl=1" "--no-profile"
run_bench next_fused_loss next "" "--fused-loss"
run_bench next_pdl_fused_loss next "pdl=1" "--fused-loss"
run_bench next_nogroup next "gemm_group=0" "--no-profile"
run_bench next_noprefetch next "epi_prefetch=0" "--no-profile"
run_bench next_alpha_dact next "alpha_from_dw2=0" "--no-profile"
run_bench next_lnreduce_main next "ln_reduce_side=0" "--no-profile"
run_bench next_dattn_dot next "dattn_from_gemm=0" "--no-profile"
run_bench next_attn_tmem_wide next "attn_tmem_compact=0"
run_bench next_defer_join next "defer_join=1"              # dW GEMMs of a block overlap the frozen LM block's backward
run_bench next_defer_join_pdl next "defer_join=1,pdl=1" "--no-profile"
run_bench next_all_off next "gemm_group=0,epi_prefetch=0,alpha_from_dw2=0,ln_reduce_side=0,dattn_from_gemm=0,attn_tmem_compact=0"
run_bench next_scalar_epilogue next_scalar ""      # same tree, -DFM_EPI_F32X2=0: attributes the packed (FFMA2) GEMM epilogues

echo "=== isolated GEMM timings + per-CTA timelines at the C2 FFW shapes: packed (next) vs scalar (next_scalar) epilogues" | tee -a "$OUT/summary.log"
for v in next next_scalar; do
  FM_B200_VARIANT=$v timeout 300 python tools/gemm_bench.py ffw1 ffw2 dact dx dw1 kv q > "$OUT/gemm_bench_$v.txt" 2>&1
  FM_B200_VARIANT=$v timeout 300 python tools/gemm_trace.py ffw1 ffw2 dact > "$OUT/gemm_trace_$v.txt" 2>&1
  echo "--- $v" | tee -a "$OUT/summary.log"; grep -v "^cta" "$OUT/gemm_bench_$v.txt" | tail -12 | tee -a "$OUT/summary.log"
  grep "^cta  0\|^==" "$OUT/gemm_trace_$v.txt" | cut -c1-400 | tee -a "$OUT/summary.log"
done

echo "=== CUPTI view of one eager step (torch.profiler: device-side kernel durations, no host launch latency inside)" | tee -a "$OUT/summary.log"
for v in "" next; do
  FM_B200_VARIANT=$v timeout 300 python tools/step_profile.py > "$OUT/step_profile_${v:-validated}.txt" 2>&1
  echo "--- variant '$v'" | tee -a "$OUT/summary.log"; grep -m1 "total CUDA kernel time" "$OUT/step_profile_${v:-validated}.txt" | tee -a "$OUT/summary.log"
done

echo "=== LayerNorm kernels in isolation: validated (two-pass) vs staging (pipelined rows)" | tee -a "$OUT/summary.log"
for v in "" next; do
  echo "--- variant '$v'" | tee -a "$OUT/summary.log"
  FM_B200_VARIANT=$v timeout 200 python tools/ln_bench.py 4096 768 2048 768 1600 768 4096 2048 2>&1 | tail -8 | tee -a "$OUT/summary.log"
done

echo "=== [next] ncu launch list" | tee -a "$OUT/summary.log"
FM_B200_VARIANT=next timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file "$OUT/ncu_launches_next.csv" python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-graph \
    > "$OUT/ncu_bench.log" 2>&1
python tools/summarize_launches.py "$OUT/ncu_launches_next.csv" 2>/dev/null | head -40 | tee -a "$OUT/summary.log"
echo "=== done" | tee -a "$OUT/summary.log"
