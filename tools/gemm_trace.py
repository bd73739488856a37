"""Per-CTA timeline of one fm_gemm_bf16 launch (clock64 stamps written by the kernel when desc.trace is set).
usage: python tools/gemm_trace.py ffw1|ffw2|dact|dw1|dx ..."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests._gpu_util import gemm  # noqa: E402
from tools.gemm_bench import CASES  # noqa: E402

DEV = "cuda"


def main():
    for name in sys.argv[1:] or ["ffw1"]:
        Mo, N, K, a_mn, b_mn, epi, out_f32, ex = CASES[name]
        A = torch.randn((K, Mo) if a_mn else (Mo, K), device=DEV).to(torch.bfloat16)
        B = torch.randn((K, N) if b_mn else (N, K), device=DEV).to(torch.bfloat16)
        aux = None
        if ex.get("aux") == "f32":
            aux = torch.randn(Mo, N, device=DEV)
        elif ex.get("aux") == "bf16":
            aux = torch.randn(Mo, N, device=DEV).to(torch.bfloat16)
        gate = torch.tensor([0.5], device=DEV)
        aux2 = None
        red = None
        kw = dict(epi=epi, out_f32=out_f32, aux=aux, aux2=aux2, red=red, out2=bool(ex.get("out2")), gate=gate)
        bn = int(os.environ.get("BN", "0"))
        gemm(A, B, a_mn, b_mn, Mo, N, K, bn=bn, **kw)
        trace = torch.zeros(148, 64, dtype=torch.int64, device=DEV)
        gemm(A, B, a_mn, b_mn, Mo, N, K, bn=bn, trace=trace, **kw)
        torch.cuda.synchronize()
        t = trace.cpu()
        print(f"== {name} M={Mo} N={N} K={K}: per-CTA clock64 deltas from kernel start (cycles)")
        for cta in (0, 1, 73, 147):
            row = t[cta]
            if row[0] == 0:
                continue
            t0 = int(row[0])
            line = [f"cta{cta:3d} end={int(row[63]) - t0:7d}"]
            for ui in range(15):
                a, b, c, d = (int(row[1 + 4 * ui + k]) for k in range(4))
                if a == 0:
                    break
                line.append(f"| u{ui}: data@{a - t0:6d} mma_issued@{b - t0:6d} epi {c - t0:6d}->{d - t0:6d} ({d - c})")
            print(" ".join(line))


if __name__ == "__main__":
    main()
