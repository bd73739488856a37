"""The CUDA-core kernels (csrc/: LayerNorm fwd / bwd / reduce with software-pipelined rows, the loss
head, the misc helpers, the fast GELU) compiled as HOST code and run thread per thread by tests/cpu_harness/simt_emu.h,
against double-precision loops (tests/cpu_harness/simt_kernels.cpp).  They were written after round 1's GPU budget was
spent; this executes their index arithmetic, shuffles, shared-memory folds and edge cases without a GPU.  It says
nothing about speed and does not cover the tcgen05 / TMA kernels (those have the unit-schedule check next door)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def emu_exe(tmp_path_factory):
    if shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_bf16.h")):
        pytest.skip("needs g++ and the CUDA headers")
    exe = str(tmp_path_factory.mktemp("simt") / "simt_kernels")
    src = os.path.join(ROOT, "tests", "cpu_harness", "simt_kernels.cpp")
    res = subprocess.run(["g++", "-std=c++20", "-O1", "-DFM_HOST_EMU", "-w", "-I", CUDA_INC, "-pthread", src, "-o", exe],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    return exe


@pytest.mark.parametrize("group", ["act", "ln_fwd", "ln_bwd", "ce", "misc"])
def test_simt_kernels_on_the_host_emulator(emu_exe, group):
    run = subprocess.run([emu_exe, group], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0 and "SIMT EMU OK" in run.stdout, run.stdout[-3000:]
