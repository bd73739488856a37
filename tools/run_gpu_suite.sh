#!/bin/bash
# Run on the GPU box (via gpurun): bring-up probe first, then each -m gpu test file in its own process so that a
# trapped kernel (sticky CUDA error) cannot mask the other files' results. Everything is tee'd into gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "=== gemm_diag" | tee gpurun_out/suite.log
timeout 300 python tools/gemm_diag.py 2>&1 | tee -a gpurun_out/suite.log
for f in tests/test_gpu_gemm.py tests/test_gpu_ops.py tests/test_gpu_modules.py "$@"; do
  echo "=== $f" | tee -a gpurun_out/suite.log
  timeout 900 python -m pytest "$f" -q -m gpu --tb=short -x 2>&1 | tail -40 | tee -a gpurun_out/suite.log
done
