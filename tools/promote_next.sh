#!/bin/bash
# Promote the staging tree to the default build once tools/validate_next.sh has passed on a B200:
#
#   bash tools/promote_next.sh            # dry run: prints what it would do
#   bash tools/promote_next.sh --apply    # does it (working tree must be clean); review `git status`, then commit
#
# * flamingo_mini_b200/csrc_next/ replaces flamingo_mini_b200/csrc/ (history kept: git mv), the old tree is removed
# * _build.py: one source tree again; "next" stays as an alias of the default build so FM_B200_VARIANT=next keeps working in
#   scripts, "next_scalar" keeps building the -DFM_EPI_F32X2=0 A/B variant from the promoted tree
# * the `first_hw_run` ordering marker is dropped from the GPU tests (they have run on hardware by then)
# Switch DEFAULTS (g_opt[] in flamingo_b200.cu) are NOT touched: set them by hand from gpurun_out/next/summary.log.
set -euo pipefail
cd "$(dirname "$0")/.."
APPLY=0; [ "${1:-}" = "--apply" ] && APPLY=1
run() { echo "+ $*"; if [ "$APPLY" = 1 ]; then "$@"; fi; }

if [ "$APPLY" = 1 ] && [ -n "$(git status --porcelain)" ]; then echo "working tree not clean" >&2; exit 1; fi
[ -d flamingo_mini_b200/csrc_next ] || { echo "no staging tree to promote" >&2; exit 1; }

run git rm -r -q flamingo_mini_b200/csrc
run git mv flamingo_mini_b200/csrc_next flamingo_mini_b200/csrc
if [ "$APPLY" = 1 ]; then
  python - <<'PY'
import re
p = "flamingo_mini_b200/_build.py"
s = open(p).read()
s = s.replace('VARIANTS = {"": "csrc", "next": "csrc_next", "next_scalar": "csrc_next"}',
              'VARIANTS = {"": "csrc", "next": "csrc", "next_scalar": "csrc"}      # promoted: one tree; "next" is an alias')
open(p, "w").write(s)
p = "flamingo_mini_b200/csrc/ptx.cuh"
s = open(p).read()
open(p, "w").write(s)                     # the emulator include path (../../tests/cpu_harness) is depth-identical: nothing to fix
p = "tests/_emu_util.py"
s = open(p).read()
s = s.replace('CSRC = os.path.join(ROOT, "flamingo_mini_b200", "csrc_next")', 'CSRC = os.path.join(ROOT, "flamingo_mini_b200", "csrc")')
open(p, "w").write(s)
import glob
for p in glob.glob("tests/cpu_harness/*") + glob.glob("tests/test_*_cpu.py"):          # harness sources include the tree by path
    s = open(p).read()
    s2 = s.replace("flamingo_mini_b200/csrc_next/", "flamingo_mini_b200/csrc/")
    if s2 != s:
        open(p, "w").write(s2)
for p in glob.glob("tests/test_gpu_*.py"):
    s = open(p).read()
    s2 = re.sub(r"^@pytest\.mark\.first_hw_run\n", "", s, flags=re.M)
    if s2 != s:
        open(p, "w").write(s2)
PY
  python -c "from flamingo_mini_b200 import _build; print(_build.build_all(force=True))"
  python -m pytest tests -x -q -m "not gpu" -k "abi or emu_selftest or group_schedule or simt or gemm_epilogues" 2>&1 | tail -3
else
  echo "+ (edit _build.py VARIANTS, tests/_emu_util.py CSRC, drop @pytest.mark.first_hw_run; rebuild; run the ABI tests)"
fi
echo "done$([ "$APPLY" = 1 ] || echo ' (dry run)'): set the switch defaults in g_opt[], update DESIGN.md section 7 / README, commit"
