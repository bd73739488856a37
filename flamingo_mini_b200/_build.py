"""Build libflamingo_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
# ONE source tree -> ONE library.  (Round 1 carried a second, not-yet-validated tree behind FM_B200_VARIANT; it was validated on a
# B200 in round 2 - profiles/r02_validate_next/ - promoted to csrc/, and the switch is gone.)
VARIANTS = {"": "csrc"}
VARIANT_FLAGS: dict = {}


def variant() -> str:
    return ""


def csrc_dir(v: str | None = None) -> str:
    return os.path.join(HERE, "csrc")


def lib_path(v: str | None = None) -> str:
    return os.path.join(HERE, "libflamingo_b200.so")


CSRC = csrc_dir("")
LIB = lib_path("")
HEADER = os.path.join(os.path.dirname(HERE), "include", "flamingo_b200.h")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libflamingo_b200.so cannot be built (there is no CPU fallback)")


def is_stale(v: str | None = None) -> bool:
    lib, src = lib_path(v), csrc_dir(v)
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = [os.path.join(src, f) for f in os.listdir(src)] + [HEADER]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, v: str | None = None) -> str:
    """Compile <csrc of the variant>/flamingo_b200.cu -> libflamingo_b200[_<variant>].so. Returns the library path."""
    lib, src = lib_path(v), csrc_dir(v)
    if not force and not is_stale(v):
        return lib
    tmp = f"{lib}.{os.getpid()}.tmp"      # per-process name: concurrent ranks may all find the library stale
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC", "-o", tmp, os.path.join(src, "flamingo_b200.cu")]
    cmd[1:1] = [f for f in os.environ.get("FM_B200_NVCC_FLAGS", "").split() if f]      # developer A/B builds only (e.g. -DFM_EPI_F32X2=1)
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    os.replace(tmp, lib)                  # atomic: readers see either the old or the new library
    if verbose:
        print(res.stderr)
    return lib


def build_all(force: bool = False) -> list[str]:
    return [build(force=force, v=v) for v in VARIANTS]


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose=True, v=(sys.argv[1] if len(sys.argv) > 1 else None)))
