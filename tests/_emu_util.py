"""Host emulation of the library (TEST INFRASTRUCTURE ONLY; nothing under flamingo_mini_b200/ imports this).

csrc/flamingo_b200.cu — the whole C ABI: launchers, tcgen05 GEMM, attention cores, LayerNorm, loss — is compiled as
plain C++ with g++ -DFM_HOST_EMU.  Kernels then run thread-per-thread on the CPU (tests/cpu_harness/simt_emu.h) against a
functional model of mbarrier / TMA / tcgen05 / TMEM (tests/cpu_harness/tc_emu.h) and a stub CUDA runtime
(tests/cpu_harness/fake_cudart.cpp).  The resulting libflamingo_b200_emu.so is built OUTSIDE the tree (temp dir) and is
only ever loaded by the CPU tests, which swap it in for the duration of one test and hand it CPU tensors.

Purpose: kernels written while no GPU was available get executed — schedules, barrier protocols (dead-locks are reported
with the kernel's own wait tags), descriptor arithmetic, epilogue indexing — before they cost GPU minutes.  It proves
nothing about performance, and nothing about hardware behaviour the model does not describe (see tc_emu.h).
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import shutil
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SRC_ROOT = os.environ.get("FM_EMU_SRC_ROOT", ROOT)          # developer override: emulate a scratch copy of csrc/ + cpu_harness/
HARNESS = os.path.join(_SRC_ROOT, "tests", "cpu_harness")
CSRC = os.path.join(_SRC_ROOT, "flamingo_mini_b200", "csrc")
CUDA_INC = "/usr/local/cuda/include"
_cached = None


def available() -> bool:
    return shutil.which("g++") is not None and os.path.exists(os.path.join(CUDA_INC, "cuda.h"))


def _sources():
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    files += [os.path.join(HARNESS, f) for f in ("simt_emu.h", "tc_emu.h", "fake_cudart.cpp")]
    files.append(os.path.join(_SRC_ROOT, "include", "flamingo_b200.h"))
    return files


def build() -> str:
    """g++ build of the emulated library, cached by a hash of its sources."""
    h = hashlib.sha256()
    for f in _sources():
        h.update(f.encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    out_dir = os.path.join(tempfile.gettempdir(), f"fm_b200_emu_{os.getuid()}", h.hexdigest()[:16])
    lib = os.path.join(out_dir, "libflamingo_b200_emu.so")
    if os.path.exists(lib):
        return lib
    os.makedirs(out_dir, exist_ok=True)
    tmp = f"{lib}.{os.getpid()}.tmp"
    cmd = ["g++", "-std=c++20", "-O2", "-DFM_HOST_EMU", "-w", "-x", "c++", "-include", os.path.join(HARNESS, "simt_emu.h"),
           "-I", CUDA_INC, "-pthread", "-fPIC", "-shared", "-Wl,-Bsymbolic",      # -Bsymbolic: bind the stub runtime, not torch's libcudart
           os.path.join(CSRC, "flamingo_b200.cu"), os.path.join(HARNESS, "fake_cudart.cpp"), "-o", tmp]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("emulation build failed:\n" + " ".join(cmd) + "\n" + res.stderr[-4000:])
    os.replace(tmp, lib)
    return lib


def load():
    """The emulated library with the package's ctypes prototypes attached ."""
    global _cached
    if _cached is None:
        from flamingo_mini_b200 import _lib
        _cached = _lib.type_library(C.CDLL(build()), staging=True)
    return _cached


class swapped_in:
    """Context manager: the package and the GPU-test helpers talk to the emulated library and accept CPU tensors."""

    def __enter__(self):
        from flamingo_mini_b200 import _lib, functional, standalone
        import tests._gpu_util as U
        self._saved = [(_lib, "_lib", _lib._lib), (functional, "_stream", functional._stream),
                       (functional, "_require_cuda", functional._require_cuda), (standalone, "_stream", standalone._stream),
                       (standalone, "_require_cuda", standalone._require_cuda), (U, "stream", U.stream)]
        _lib._lib = load()
        for mod in (functional, standalone):
            mod._stream = lambda: None
            mod._require_cuda = lambda t, what: None
        U.stream = lambda: None
        return _lib._lib

    def __exit__(self, *exc):
        for obj, name, val in self._saved:
            setattr(obj, name, val)
        return False
