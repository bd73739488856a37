"""The emulator's own negative tests (tests/cpu_harness/tc_selftest.cpp): protocol mistakes must be CAUGHT — data that was
not waited for is not there yet (lazy completion), wrong TMEM lanes / unallocated columns / misaligned swizzled tiles abort
with a message, a transaction-count mismatch is reported as a dead-lock with the kernel's wait tag."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    if shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA_INC, "cuda.h")):
        pytest.skip("needs g++ and the CUDA headers")
    out = str(tmp_path_factory.mktemp("tcself") / "tc_selftest")
    res = subprocess.run(["g++", "-std=c++20", "-O1", "-w", "-I", CUDA_INC, "-pthread",
                          os.path.join(ROOT, "tests", "cpu_harness", "tc_selftest.cpp"), "-o", out], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    return out


@pytest.mark.parametrize("case", ["ok", "no_tma_wait", "no_mma_wait"])
def test_model_runs_and_completes_lazily(exe, case):
    run = subprocess.run([exe, case], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0 and f"SELFTEST {case} OK" in run.stdout, run.stdout + run.stderr


@pytest.mark.parametrize("case,message", [("wrong_lanes", "may only touch TMEM lanes"), ("tx_mismatch", "dead-lock"),
                                          ("misaligned", "not 1024-byte aligned"), ("unallocated", "are not allocated")])
def test_model_catches_protocol_mistakes(exe, case, message):
    run = subprocess.run([exe, case], capture_output=True, text=True, timeout=120, env=dict(os.environ, FM_EMU_TIMEOUT_S="2"))
    assert run.returncode != 0 and message in run.stderr, run.stdout + run.stderr
    if case == "tx_mismatch":
        assert "tag 0x777" in run.stderr
