#!/bin/bash
# ncu --set full captures of the kernels that carry the C2 step, IN SITU (inside one eager training step, cold caches as in the
# real step), one launch each.  ONE gpurun call on ONE GPU (ncu replays every captured launch ~40 times):
#
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash tools/ncu_top_kernels.sh'
#
# Reports land in gpurun_out/ncu/*.ncu-rep; read them HERE afterwards (no GPU needed):
#
#   for f in gpurun_out/ncu/*.ncu-rep; do echo "== $f"; ncu -i $f --page raw --csv | python tools/ncu_pick.py; done
#   ncu -i gpurun_out/ncu/act_gemm.ncu-rep --page source --csv > gpurun_out/ncu/act_gemm_source.csv     # per-line stalls (-lineinfo)
#
# and summarise into profiles/rNN_ncu_*.md (tensor-pipe utilisation, issue-active, DRAM bytes vs the algorithmic bytes of
# DESIGN.md section 3); put dram read+write bytes into bench.NCU_TRAFFIC_BYTES.  Numbers printed by the run under ncu are never
# bench values.
cd "$(dirname "$0")/.."
OUT=gpurun_out/ncu
mkdir -p "$OUT"
CMD="python bench.py --steps 1 --warmup 1 --no-graph --no-profile --no-cpu-baseline"
# name | regex on the MANGLED kernel name | launches of that kernel to skip first (warm-up step + first instance of the timed step)
while IFS='|' read -r name regex skip; do
  [ -z "$name" ] && continue
  echo "=== $name  ($regex, skip $skip)"
  timeout 420 ncu --set full --clock-control none --import-source on --kernel-name-base mangled \
      -k "regex:$regex" -s "$skip" -c 1 \
      -f -o "$OUT/$name" $CMD > "$OUT/$name.log" 2>&1
  tail -2 "$OUT/$name.log"
done <<'EOF'
act_gemm|_ZN2fm14gemm_tc_kernelILi256ELb0ELb0ELi1E|14
dact_gemm|_ZN2fm14gemm_tc_kernelILi256ELb0ELb1ELi3E|14
resid_gemm|_ZN2fm14gemm_tc_kernelILi192ELb0ELb0ELi2E|26
dw_gemm|_ZN2fm14gemm_tc_kernelILi128ELb1ELb1ELi0E|40
ln_bwd_dx|_ZN2fm18ln_bwd_dx_w_kernel|50
xattn_core_bwd|_ZN2fm24xattn_core_bwd_tc_kernel|14
EOF
ls -la "$OUT"
