"""Isolated timing of the LayerNorm kernels through the C ABI (CUDA events, L2 flushed).  usage: python tools/ln_bench.py [rows D]..."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flamingo_mini_b200 import _lib  # noqa: E402
from flamingo_mini_b200._lib import check  # noqa: E402
from tests._gpu_util import ptr, stream  # noqa: E402

DEV = "cuda"


def main():
    lib = _lib.load()
    args = [int(a) for a in sys.argv[1:]] or [4096, 768, 4096, 2048]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    for rows, D in zip(args[::2], args[1::2]):
        x = torch.randn(rows, D, device=DEV)
        gamma, beta = torch.ones(D, device=DEV), torch.zeros(D, device=DEV)
        out = torch.empty(rows, D, device=DEV, dtype=torch.bfloat16)
        mean, rstd = torch.empty(rows, device=DEV), torch.empty(rows, device=DEV)
        dy = torch.randn(rows, D, device=DEV).to(torch.bfloat16)
        dres = torch.randn(rows, D, device=DEV).to(torch.bfloat16)
        dx = torch.empty(rows, D, device=DEV, dtype=torch.bfloat16)
        dg, db = torch.empty(D, device=DEV), torch.empty(D, device=DEV)
        part = torch.empty(lib.fm_layernorm_bwd_scratch_bytes(D), dtype=torch.uint8, device=DEV)

        def fwd():
            check(lib.fm_layernorm_fwd(ptr(x), 1, ptr(gamma), ptr(beta), ptr(out), 0, ptr(mean), ptr(rstd), rows, D, stream()))

        def bwd():
            check(lib.fm_layernorm_bwd(ptr(dy), ptr(x), 1, ptr(gamma), ptr(mean), ptr(rstd), ptr(dres), 0, ptr(dx), 0, ptr(dg), ptr(db),
                                       ptr(part), rows, D, stream()))
        for name, fn, nbytes in (("ln_fwd", fwd, rows * D * 6), ("ln_bwd(+reduce)", bwd, rows * D * 10)):
            fn(); torch.cuda.synchronize()
            ts = []
            for cold in (True, False):
                for _ in range(5):
                    if cold:
                        flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda._sleep(600_000)       # GPU-side head start: the launch is queued before the GPU reaches e0
                    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                    ts.append((cold, e0.elapsed_time(e1)))
            c = sorted(t for k, t in ts if k)[2]
            w = sorted(t for k, t in ts if not k)[2]
            print(f"{name:16s} rows={rows} D={D}: cold {c*1e3:7.1f} us ({nbytes/c/1e6:6.0f} GB/s)  warm {w*1e3:7.1f} us ({nbytes/w/1e6:6.0f} GB/s)")


if __name__ == "__main__":
    main()
