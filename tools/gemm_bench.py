"""Isolated timing of fm_gemm_bf16 on the hot-path shapes (CUDA events, L2 flushed between launches).
usage: python tools/gemm_bench.py [case ...]   cases: ffw1 ffw2 dact dw1 dw2 dwq kv q   (default: all)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests._gpu_util import gemm  # noqa: E402

DEV = "cuda"
M, D, FF, I = 4096, 768, 3072, 512
CASES = {
    #        M    N    K   a_mn b_mn epi out_f32 extra
    "ffw1": (M, FF, D, 0, 0, 1, False, dict(out2=True)),
    "ffw2": (M, D, FF, 0, 0, 2, False, dict(aux="f32")),
    "dact": (M, FF, D, 0, 1, 3, False, dict(aux="bf16")),
    "dx":   (M, D, FF, 0, 1, 0, False, {}),
    "dw1":  (FF, D, M, 1, 1, 0, True, dict(flags=True)),
    "dw2":  (D, FF, M, 1, 1, 0, True, dict(flags=True)),
    "dwq":  (I, D, M, 1, 1, 0, True, dict(flags=True)),
    "dwq_nosplit": (I, D, M, 1, 1, 0, True, {}),
    "kv":   (3648, 1024, D, 0, 0, 0, False, {}),
    "q":    (2048, I, D, 0, 0, 0, False, {}),
    "big":  (8192, 8192, 8192, 0, 0, 0, False, {}),
}


def main():
    names = sys.argv[1:] or list(CASES)
    reps = int(os.environ.get("REPS", "10"))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    for name in names:
        Mo, N, K, a_mn, b_mn, epi, out_f32, ex = CASES[name]
        A = torch.randn((K, Mo) if a_mn else (Mo, K), device=DEV).to(torch.bfloat16)
        B = torch.randn((K, N) if b_mn else (N, K), device=DEV).to(torch.bfloat16)
        aux = None
        if ex.get("aux") == "f32":
            aux = torch.randn(Mo, N, device=DEV)
        elif ex.get("aux") == "bf16":
            aux = torch.randn(Mo, N, device=DEV).to(torch.bfloat16)
        flags = torch.zeros(16384, dtype=torch.int32, device=DEV) if ex.get("flags") else None
        gate = torch.tensor([0.5], device=DEV)
        aux2 = None
        red = None
        kw = dict(epi=epi, out_f32=out_f32, aux=aux, aux2=aux2, red=red, out2=bool(ex.get("out2")), gate=gate, flags=flags)
        for bn in ([0] if os.environ.get("BN") is None else [int(os.environ["BN"])]):
            gemm(A, B, a_mn, b_mn, Mo, N, K, bn=bn, **kw)
            torch.cuda.synchronize()
            ts = []
            for _ in range(reps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda._sleep(600_000)      # ~0.3 ms of GPU spin: the launch below is enqueued before the GPU reaches e0, so
                e0.record()                     # the interval holds the kernel, not the host's launch path (ctypes + tensor maps)
                gemm(A, B, a_mn, b_mn, Mo, N, K, bn=bn, **kw)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            med = ts[len(ts) // 2]
            print(f"{name:12s} M={Mo} N={N} K={K} bn={bn}: median {med*1e3:8.1f} us  min {ts[0]*1e3:8.1f} us  "
                  f"{2.0*Mo*N*K/med/1e9:7.1f} TFLOP/s (median)", flush=True)


if __name__ == "__main__":
    main()
