#!/bin/bash
# ONE 1-GPU gpurun call of round 2 on the promoted library (everything lands in gpurun_out/r02/):
#   /usr/local/graft/bin/gpurun --timeout 2700 -- 'bash tools/gpu_call_r02.sh'
# 1. pytest -m gpu (whole suite) + smoke()      2. bench C2 (graph-event profiler -> roofline, xattn key)
# 3. model-level parity report + reference-eager-on-B200 timing      4. compute-sanitizer memcheck / racecheck logs
# 5. bench C3 / C4 / C5 at their per-GPU batch on one GPU            6. ncu --set full in situ of the six kernels that carry the step
cd "$(dirname "$0")/.."
OUT=gpurun_out/r02
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/smi.txt" 2>&1
echo "=== pytest -m gpu" | tee "$OUT/summary.log"
timeout 1200 python -m pytest tests -q -m gpu --tb=short -s 2>&1 | grep -v "^$" | tail -45 | tee -a "$OUT/summary.log"
echo "=== smoke" | tee -a "$OUT/summary.log"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee -a "$OUT/summary.log"
echo "=== bench c2" | tee -a "$OUT/summary.log"
timeout 600 python bench.py --steps 30 --warmup 5 > "$OUT/bench_c2.json" 2> "$OUT/bench_c2.err"
python - "$OUT/bench_c2.json" <<'PY' | tee -a "$OUT/summary.log"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(f"  {d['ms_per_step']:.3f} ms/step {d['value']:.1f} samples/s e2e {d['e2e']['value']:.1f} gemm frac {r.get('frac')} lib ms {r.get('library_kernel_ms_per_step')} "
          f"profiled step {r.get('step_ms_under_profiler_events')} xattn {d.get('xattn')} cpu {d.get('cpu_baseline')}")
    for i in (r.get("instantiations") or []):
        print("   ", i["tag"], i["launches_per_step"], round(i["avg_launch_ms"] * 1e3, 1), "us", round(i["achieved"]), "TF", round(i["frac"], 3))
    for k, v in list((d.get("kernels") or {}).items())[:30]:
        print("    k", k, v["launches"], round(v["ms_per_step"], 3))
except Exception as e:
    print("  FAILED:", e)
PY
echo "=== parity report" | tee -a "$OUT/summary.log"
timeout 900 python tools/parity_report.py --out "$OUT/r02_parity.md" 2>&1 | tail -30 | tee -a "$OUT/summary.log"
echo "=== CUPTI step profile" | tee -a "$OUT/summary.log"
timeout 300 python tools/step_profile.py > "$OUT/step_profile.txt" 2>&1; grep -m1 "total CUDA kernel time" "$OUT/step_profile.txt" | tee -a "$OUT/summary.log"
echo "=== compute-sanitizer" | tee -a "$OUT/summary.log"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python __graft_entry__.py smoke > "$OUT/memcheck_smoke.log" 2>&1
echo "  memcheck smoke: exit $? ; $(tail -1 "$OUT/memcheck_smoke.log")" | tee -a "$OUT/summary.log"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python __graft_entry__.py smoke > "$OUT/racecheck_smoke.log" 2>&1
echo "  racecheck smoke: exit $? ; $(tail -1 "$OUT/racecheck_smoke.log")" | tee -a "$OUT/summary.log"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python bench.py --steps 1 --warmup 1 --no-graph --no-profile --no-cpu-baseline > "$OUT/memcheck_c2_step.log" 2>&1
echo "  memcheck one C2 step (eager): exit $? ; $(tail -1 "$OUT/memcheck_c2_step.log")" | tee -a "$OUT/summary.log"
for wl in c3 c4 c5; do
  echo "=== bench $wl (1 GPU, per-GPU batch of the config)" | tee -a "$OUT/summary.log"
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > "$OUT/bench_$wl.json" 2> "$OUT/bench_$wl.err"
  python - "$OUT/bench_$wl.json" <<'PY' | tee -a "$OUT/summary.log"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(f"  {d['ms_per_step']:.3f} ms/step {d['value']:.1f} samples/s gemm frac {r.get('frac')} lib ms {r.get('library_kernel_ms_per_step')} xattn {(d.get('xattn') or {}).get('frac')}")
except Exception as e:
    print("  FAILED:", e)
PY
done
echo "=== ncu in situ" | tee -a "$OUT/summary.log"
bash tools/ncu_top_kernels.sh 2>&1 | tail -20 | tee -a "$OUT/summary.log"
mkdir -p "$OUT/ncu"
for f in gpurun_out/ncu/*.ncu-rep; do
  n=$(basename "$f" .ncu-rep)
  ncu -i "$f" --page raw --csv 2>/dev/null | python tools/ncu_pick.py > "$OUT/ncu/$n.txt" 2>&1
done
du -sh gpurun_out | tee -a "$OUT/summary.log"
echo "=== done" | tee -a "$OUT/summary.log"
