"""Stand-alone bring-up probe for the tcgen05 GEMM (run on the GPU box before the pytest suite).
Prints one line per (layout, shape) case and, on a mismatch, a per-k rank-1 probe that shows which K slices of a
64-wide block land in the wrong place (descriptor / swizzle bugs)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests._gpu_util import gemm, logical, rel_err  # noqa: E402
from flamingo_mini_b200 import _lib  # noqa: E402


def make(rows, K, mn, dev, gen):
    X = torch.randn((K, rows) if mn else (rows, K), device=dev, generator=gen).to(torch.bfloat16)
    return X


def probe(a_mn, b_mn, bn, dev, gen):
    M, N, K = 128, bn, 64
    A = make(M, K, a_mn, dev, gen)
    bad = []
    for k0 in range(K):
        B = torch.zeros((K, N) if b_mn else (N, K), device=dev, dtype=torch.bfloat16)
        if b_mn:
            B[k0, :] = 1.0
        else:
            B[:, k0] = 1.0
        out = gemm(A, B, a_mn, b_mn, M, N, K, out_f32=True, bn=bn)
        ref = logical(A, a_mn)[:, k0:k0 + 1].expand(M, N)
        if not torch.allclose(out, ref, atol=1e-2, rtol=1e-2):
            bad.append(k0)
    print(f"    k-probe a_mn={a_mn} b_mn={b_mn} bn={bn}: bad k = {bad}")


def main():
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(0)
    lib = _lib.load()
    ok_all = True
    shapes = [(128, 64, 64), (128, 128, 64), (128, 256, 64), (128, 192, 128), (256, 256, 256), (384, 512, 768),
              (1000, 520, 200), (4096, 3072, 768)]
    for (a_mn, b_mn) in [(0, 0), (0, 1), (1, 1)]:
        for (M, N, K) in shapes:
            for bn in ([64, 128, 192, 256] if (M, N, K) == (256, 256, 256) else [0]):
                A, B = make(M, K, a_mn, dev, gen), make(N, K, b_mn, dev, gen)
                try:
                    out = gemm(A, B, a_mn, b_mn, M, N, K, out_f32=True, bn=bn)
                    torch.cuda.synchronize()
                except Exception as e:
                    print(f"a_mn={a_mn} b_mn={b_mn} M={M} N={N} K={K} bn={bn}: EXCEPTION {e}")
                    print("device error word:", hex(lib.fm_device_error()))
                    return 1
                ref = logical(A, a_mn) @ logical(B, b_mn).t()
                err = rel_err(out, ref)
                nan = int(torch.isnan(out).sum())
                good = err < 2e-3 and nan == 0
                ok_all &= good
                print(f"a_mn={a_mn} b_mn={b_mn} M={M} N={N} K={K} bn={bn}: rel_err={err:.3e} nan={nan} {'ok' if good else 'FAIL'}")
                if not good and K >= 64:
                    for pb in (64, 128):
                        probe(a_mn, b_mn, pb, dev, gen)
                    break
    print("device error word:", hex(lib.fm_device_error()))
    print("GEMM DIAG", "PASS" if ok_all else "FAIL")
    return 0 if ok_all else 2


if __name__ == "__main__":
    sys.exit(main())
