"""The persistent GEMM's unit schedule (grouped launches + serial split-K) checked on the host: the
kernel's own locate_unit / tile_coords are compiled as host code (tests/cpu_harness/group_schedule.cu) and every
(problem, tile, K range) must be produced exactly once.  No GPU needed."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="nvcc not available")
def test_group_schedule_covers_every_tile_once(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "group_schedule")
    src = os.path.join(ROOT, "tests", "cpu_harness", "group_schedule.cu")
    res = subprocess.run([nvcc, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-w", "-o", exe, src],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0 and "GROUP SCHEDULE OK" in run.stdout, run.stdout[-2000:]
