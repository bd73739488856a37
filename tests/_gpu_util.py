"""Shared helpers for the -m gpu tests: thin ctypes-level wrappers so kernels are exercised through the C ABI."""
import ctypes as C

import torch

from flamingo_mini_b200 import _lib
from flamingo_mini_b200._lib import GemmDesc, check


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def gemm(A, B, a_mn, b_mn, M, N, K, epi=0, out_f32=False, aux=None, aux2=None, out2=False, gate=None, scale=1.0, act=0,
         bias=None, red=None, bn=0, splits=0, flags=None, trace=None):
    """D[m,n] = sum_k A(m,k) B(n,k) through fm_gemm_bf16. A/B are 2-D bf16 tensors in their stored layout.
    splits < 0 = parallel split-K: the output is handed over zeroed (every K range reduce-adds into it)."""
    lib = _lib.load()
    out = torch.full((M, N), 0.0 if splits < 0 else float("nan"), dtype=torch.float32 if out_f32 else torch.bfloat16, device=A.device)
    o2 = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=A.device) if out2 else None
    d = GemmDesc(M=M, N=N, K=K, A=ptr(A), lda=A.stride(0), a_mn=int(a_mn), B=ptr(B), ldb=B.stride(0), b_mn=int(b_mn),
                 epi=epi, out=ptr(out), ldo=N, out_f32=int(out_f32), out2=ptr(o2), ldo2=N,
                 aux=ptr(aux), ldaux=(aux.stride(0) if aux is not None else 0),
                 aux_f32=int(aux is not None and aux.dtype == torch.float32),
                 aux2=ptr(aux2), ldaux2=(aux2.stride(0) if aux2 is not None else 0),
                 col_bias=ptr(bias), gate=ptr(gate), red_out=ptr(red), scale=scale, act=act, bn=bn, splits=splits, splitk_flags=ptr(flags), trace=ptr(trace))
    check(lib.fm_gemm_bf16(d, stream()), "fm_gemm_bf16")
    return (out, o2) if out2 else out


def gemm_group(problems):
    """problems: list of dicts(A, B, a_mn, b_mn, M, N, K, out_f32, scale, gate) -> list of outputs, through
    fm_gemm_bf16_group (ONE persistent launch; n sequential launches with the gemm_group switch off)."""
    lib = _lib.load()
    n = len(problems)
    descs = (GemmDesc * n)()
    outs = []
    for i, q in enumerate(problems):
        A, B, M, N, K = q["A"], q["B"], q["M"], q["N"], q["K"]
        out = torch.full((M, N), 0.0 if q.get("splits", 0) < 0 else float("nan"), dtype=torch.float32 if q.get("out_f32") else torch.bfloat16,
                         device=A.device)
        outs.append(out)
        descs[i] = GemmDesc(M=M, N=N, K=K, A=ptr(A), lda=A.stride(0), a_mn=int(q["a_mn"]), B=ptr(B), ldb=B.stride(0),
                            b_mn=int(q["b_mn"]), epi=0, out=ptr(out), ldo=N, out_f32=int(bool(q.get("out_f32"))),
                            gate=ptr(q.get("gate")), scale=q.get("scale", 1.0), bn=q.get("bn", 0), splits=q.get("splits", 0))
    check(lib.fm_gemm_bf16_group(descs, n, stream()), "fm_gemm_bf16_group")
    return outs


def set_option(name, value):
    """fm_set_option by name; returns False when the loaded library does not know the switch."""
    return _lib.load().fm_set_option(_lib.OPTION_KEYS[name], int(value)) == 0


def logical(X, mn):
    """stored tensor -> logical [rows(M or N), K] fp32 matrix"""
    return (X.t() if mn else X).float()


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()
