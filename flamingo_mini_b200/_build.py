"""Build libflamingo_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
# Two source trees with the SAME C ABI (include/flamingo_b200.h):
#   csrc/       every kernel in it has passed `pytest -m gpu` on a B200 -> libflamingo_b200.so (what is loaded by default)
#   csrc_next/  staging tree: kernels written without GPU access, to be validated with tools/validate_next.sh and then
#               promoted (git mv csrc_next csrc) -> libflamingo_b200_next.so, loaded ONLY when FM_B200_VARIANT=next
#   next_scalar: the staging tree compiled with -DFM_EPI_F32X2=0 (GEMM epilogue arithmetic with scalar fp32 instead of the
#               packed FFMA2 / FMUL2 forms): exists only so that ONE GPU call can attribute the gain of the packed epilogues
VARIANTS = {"": "csrc", "next": "csrc_next", "next_scalar": "csrc_next"}
VARIANT_FLAGS = {"next_scalar": ["-DFM_EPI_F32X2=0"]}


def variant() -> str:
    v = os.environ.get("FM_B200_VARIANT", "")
    if v not in VARIANTS:
        raise RuntimeError(f"FM_B200_VARIANT={v!r}: expected one of {sorted(VARIANTS)}")
    return v


def csrc_dir(v: str | None = None) -> str:
    return os.path.join(HERE, VARIANTS[variant() if v is None else v])


def lib_path(v: str | None = None) -> str:
    v = variant() if v is None else v
    return os.path.join(HERE, f"libflamingo_b200{'_' + v if v else ''}.so")


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
    cmd[1:1] = VARIANT_FLAGS.get(variant() if v is None else v, [])
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
