"""Stand-alone (inference) forwards of the sub-modules the reference also exposes as nn.Modules:
``FeedForward`` (utils.py:31-50), ``MaskedCrossAttention.forward`` (gated_cross_attention.py:42-131) and
``PerceiverAttentionLayer.forward`` (perceiver_resampler.py:32-96).

Inside ``PerceiverResampler`` / ``GatedCrossAttentionBlock`` these run fused (fm_resampler_* / fm_xattn_*), which is
the training path.  Called on their own they are composed here from the library's primitives — fm_layernorm_fwd,
fm_gemm_bf16 and the attention cores (fm_{xattn,resampler}_core_{fwd,bwd}).  All three are differentiable on their own
(FeedForward's backward is four more GEMMs and a LayerNorm backward; the attention modules use the core backward entry
points); with cached keys/values they are inference-only — a tensor that requires grad is then rejected rather than silently
detached.  No CPU / eager fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import ACT_IDS, FlamingoB200Error, GemmDesc, check
from .functional import _ptr, _require_cuda, _stream, text_time_of

EPI_STORE, EPI_ACT = 0, 1


def _inference_only(what: str, *tensors) -> None:
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise FlamingoB200Error(f"{what}: the stand-alone forward is inference-only (gradients flow through the fused "
                                "PerceiverResampler / GatedCrossAttentionBlock path); call it under torch.no_grad()")


def _need(name: str) -> None:
    if not _lib.has(name):
        raise FlamingoB200Error(f"{name} is not exported by the loaded libflamingo_b200.so (stale build?)")


def _layernorm(x2d: torch.Tensor, norm: torch.nn.LayerNorm) -> torch.Tensor:
    rows, D = x2d.shape
    out = torch.empty((rows, D), dtype=torch.bfloat16, device=x2d.device)
    w, b = norm.weight.detach().float().contiguous(), norm.bias.detach().float().contiguous()
    check(_lib.load().fm_layernorm_fwd(_ptr(x2d), int(x2d.dtype == torch.float32), _ptr(w), _ptr(b), _ptr(out), 0, None, None,
                                       rows, D, _stream()), "fm_layernorm_fwd")
    return out


def _ln_fwd_stats(x2d: torch.Tensor, norm: torch.nn.LayerNorm):
    """LayerNorm forward that also returns what its backward needs: (bf16 output, fp32 gamma, mean, rstd)."""
    rows, D = x2d.shape
    out = torch.empty((rows, D), dtype=torch.bfloat16, device=x2d.device)
    g32, b32 = norm.weight.detach().float().contiguous(), norm.bias.detach().float().contiguous()
    mean, rstd = torch.empty(rows, dtype=torch.float32, device=x2d.device), torch.empty(rows, dtype=torch.float32, device=x2d.device)
    check(_lib.load().fm_layernorm_fwd(_ptr(x2d), int(x2d.dtype == torch.float32), _ptr(g32), _ptr(b32), _ptr(out), 0, _ptr(mean),
                                       _ptr(rstd), rows, D, _stream()), "fm_layernorm_fwd")
    return out, g32, mean, rstd


def _ln_bwd(dy_bf16: torch.Tensor, x2d: torch.Tensor, g32, mean, rstd):
    """-> (dx in x's dtype, dgamma fp32, dbeta fp32) through fm_layernorm_bwd."""
    lib = _lib.load()
    rows, D = x2d.shape
    dx = torch.empty_like(x2d)
    dg, db = torch.empty(D, dtype=torch.float32, device=x2d.device), torch.empty(D, dtype=torch.float32, device=x2d.device)
    part = torch.empty(lib.fm_layernorm_bwd_scratch_bytes(D), dtype=torch.uint8, device=x2d.device)
    check(lib.fm_layernorm_bwd(_ptr(dy_bf16), _ptr(x2d), int(x2d.dtype == torch.float32), _ptr(g32), _ptr(mean), _ptr(rstd), None, 0,
                               _ptr(dx), int(dx.dtype == torch.float32), _ptr(dg), _ptr(db), _ptr(part), rows, D, _stream()),
          "fm_layernorm_bwd")
    return dx, dg, db


def _bf16(w: torch.Tensor) -> torch.Tensor:
    return w.detach().to(torch.bfloat16).contiguous()


def _linear(x2d: torch.Tensor, weight: torch.Tensor, epi: int = EPI_STORE, scale: float = 1.0, act: str = "gelu",
            out_dtype=torch.bfloat16) -> torch.Tensor:
    """y = x W^T through fm_gemm_bf16 (both operands K-major); x2d bf16 [M, K], weight [N, K] (any float dtype)."""
    M, K = x2d.shape
    N = weight.shape[0]
    w = weight.detach().to(torch.bfloat16).contiguous()
    out = torch.empty((M, N), dtype=out_dtype, device=x2d.device)
    d = GemmDesc(M=M, N=N, K=K, A=_ptr(x2d), lda=K, a_mn=0, B=_ptr(w), ldb=K, b_mn=0, epi=epi, out=_ptr(out), ldo=N,
                 out_f32=int(out_dtype == torch.float32), scale=scale, act=ACT_IDS[act])
    check(_lib.load().fm_gemm_bf16(C.byref(d), _stream()), "fm_gemm_bf16")
    return out


def _as_rows(x: torch.Tensor, what: str) -> torch.Tensor:
    _require_cuda(x, what)
    if x.dtype not in (torch.bfloat16, torch.float32):
        raise FlamingoB200Error(f"{what}: unsupported dtype {x.dtype}: use bfloat16 or float32")
    return x.reshape(-1, x.shape[-1]).contiguous()


EPI_DACT = 3


def _gemm(M, N, K, A, lda, a_mn, B, ldb, b_mn, out, epi=EPI_STORE, **kw) -> None:
    d = GemmDesc(M=M, N=N, K=K, A=_ptr(A), lda=lda, a_mn=a_mn, B=_ptr(B), ldb=ldb, b_mn=b_mn, epi=epi, out=_ptr(out),
                 ldo=N, out_f32=int(out.dtype == torch.float32), scale=1.0, **kw)
    check(_lib.load().fm_gemm_bf16(C.byref(d), _stream()), "fm_gemm_bf16")


class _FeedForwardFn(torch.autograd.Function):
    """FeedForward (utils.py:31-50) with gradients, composed from the library's primitives only — the same kernels the fused
    modules use, in the same three operand layouts (y = x W^T, dx = dy W, dW = dy^T x: nothing is transposed in HBM):
        forward   xn = LN(x) ; h = act(xn W1^T) (epilogue also saves act') ; out = h W2^T
        backward  dh = (dout W2) * act'  (DACT epilogue) ; dW2 = dout^T h ; dW1 = dh^T xn ; dxn = dh W1 ; LN backward."""

    @staticmethod
    def forward(ctx, x2, ln_w, ln_b, w1, w2, act):
        lib = _lib.load()
        M, D = x2.shape
        Fi = w1.shape[0]
        dev = x2.device
        g32, b32 = ln_w.detach().float().contiguous(), ln_b.detach().float().contiguous()
        w1b, w2b = w1.detach().to(torch.bfloat16).contiguous(), w2.detach().to(torch.bfloat16).contiguous()
        xn = torch.empty((M, D), dtype=torch.bfloat16, device=dev)
        mean, rstd = torch.empty(M, dtype=torch.float32, device=dev), torch.empty(M, dtype=torch.float32, device=dev)
        check(lib.fm_layernorm_fwd(_ptr(x2), int(x2.dtype == torch.float32), _ptr(g32), _ptr(b32), _ptr(xn), 0, _ptr(mean), _ptr(rstd),
                                   M, D, _stream()), "fm_layernorm_fwd")
        h = torch.empty((M, Fi), dtype=torch.bfloat16, device=dev)
        dact = torch.empty((M, Fi), dtype=torch.bfloat16, device=dev)
        _gemm(M, Fi, D, xn, D, 0, w1b, D, 0, h, epi=EPI_ACT, act=ACT_IDS[act], out2=_ptr(dact), ldo2=Fi)
        out = torch.empty((M, D), dtype=x2.dtype, device=dev)
        _gemm(M, D, Fi, h, Fi, 0, w2b, Fi, 0, out)
        ctx.save_for_backward(x2, g32, mean, rstd, xn, h, dact, w1b, w2b)
        ctx.act, ctx.param_dtypes = act, (ln_w.dtype, ln_b.dtype, w1.dtype, w2.dtype)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        x2, g32, mean, rstd, xn, h, dact, w1b, w2b = ctx.saved_tensors
        M, D = x2.shape
        Fi = w1b.shape[0]
        dev = x2.device
        do = dout.to(torch.bfloat16).contiguous()
        dh = torch.empty((M, Fi), dtype=torch.bfloat16, device=dev)
        _gemm(M, Fi, D, do, D, 0, w2b, Fi, 1, dh, epi=EPI_DACT, act=ACT_IDS[ctx.act], aux=_ptr(dact), ldaux=Fi)   # B = W2 stored [D, Fi]
        dw2 = torch.empty((D, Fi), dtype=torch.float32, device=dev)
        _gemm(D, Fi, M, do, D, 1, h, Fi, 1, dw2)                                                                  # dout^T h
        dw1 = torch.empty((Fi, D), dtype=torch.float32, device=dev)
        _gemm(Fi, D, M, dh, Fi, 1, xn, D, 1, dw1)                                                                 # dh^T xn
        dxn = torch.empty((M, D), dtype=torch.bfloat16, device=dev)
        _gemm(M, D, Fi, dh, Fi, 0, w1b, D, 1, dxn)                                                                # dh W1
        dx = torch.empty((M, D), dtype=x2.dtype, device=dev)
        dg, db = torch.empty(D, dtype=torch.float32, device=dev), torch.empty(D, dtype=torch.float32, device=dev)
        part = torch.empty(lib.fm_layernorm_bwd_scratch_bytes(D), dtype=torch.uint8, device=dev)
        check(lib.fm_layernorm_bwd(_ptr(dxn), _ptr(x2), int(x2.dtype == torch.float32), _ptr(g32), _ptr(mean), _ptr(rstd), None, 0,
                                   _ptr(dx), int(dx.dtype == torch.float32), _ptr(dg), _ptr(db), _ptr(part), M, D, _stream()),
              "fm_layernorm_bwd")
        t = ctx.param_dtypes
        return dx, dg.to(t[0]), db.to(t[1]), dw1.to(t[2]), dw2.to(t[3]), None


def feed_forward(ff, x: torch.Tensor) -> torch.Tensor:
    """LayerNorm -> Linear -> act -> Linear (utils.py:45-50); x (..., dim) -> same shape and dtype.  With grad mode on and
    anything requiring grad, the autograd path above runs (it also saves act'); otherwise the lean inference composition."""
    x2 = _as_rows(x, "FeedForward input")
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in ff.parameters())):
        for p in ff.parameters():
            _require_cuda(p, "FeedForward parameter")
        out = _FeedForwardFn.apply(x2, ff[0].weight, ff[0].bias, ff[1].weight, ff[3].weight, ff.act)
        return out.view(x.shape)
    h = _linear(_layernorm(x2, ff[0]), ff[1].weight, epi=EPI_ACT, act=ff.act)
    out = _linear(h, ff[3].weight, out_dtype=x.dtype)
    return out.view(x.shape)


class _MaskedCrossAttentionFn(torch.autograd.Function):
    """MaskedCrossAttention (gated_cross_attention.py:42-131) with gradients, from primitives + the two attention-core entry
    points.  q is scaled by dim_head^-0.5 in the projection epilogue; the core backward returns dq with respect
    to the UN-scaled projection (it applies the same factor), so dWq = dq^T yn and dyn = dq Wq need no further scaling."""

    @staticmethod
    def forward(ctx, y2, vis2, tt, norm_w, norm_b, wq, wkv, wout, dims, norm):
        B, S, n_media, H, scale = dims
        M, D = y2.shape
        I = 64 * H
        dev = y2.device
        yn, g32, mean, rstd = _ln_fwd_stats(y2, norm)
        wqb, wkvb, woutb = _bf16(wq), _bf16(wkv), _bf16(wout)
        q = torch.empty((M, I), dtype=torch.bfloat16, device=dev)
        d = GemmDesc(M=M, N=I, K=D, A=_ptr(yn), lda=D, a_mn=0, B=_ptr(wqb), ldb=D, b_mn=0, epi=EPI_STORE, out=_ptr(q), ldo=I, out_f32=0, scale=scale)
        check(_lib.load().fm_gemm_bf16(C.byref(d), _stream()), "fm_gemm_bf16")
        V, Dv = vis2.shape
        kv = torch.empty((V, 2 * I), dtype=torch.bfloat16, device=dev)
        _gemm(V, 2 * I, Dv, vis2, Dv, 0, wkvb, Dv, 0, kv)
        o = torch.empty((M, I), dtype=torch.bfloat16, device=dev)
        check(_lib.load().fm_xattn_core_fwd(_ptr(q), _ptr(kv), _ptr(tt), _ptr(o), B, S, n_media, H, _stream()), "fm_xattn_core_fwd")
        out = torch.empty((M, D), dtype=y2.dtype, device=dev)
        _gemm(M, D, I, o, I, 0, woutb, I, 0, out)
        ctx.save_for_backward(y2, vis2, tt, g32, mean, rstd, yn, q, kv, o, wqb, wkvb, woutb)
        ctx.dims, ctx.dtypes = dims, (norm_w.dtype, norm_b.dtype, wq.dtype, wkv.dtype, wout.dtype)
        ctx.mark_non_differentiable(kv)
        return out, kv

    @staticmethod
    def backward(ctx, dout, _dkv_unused):
        y2, vis2, tt, g32, mean, rstd, yn, q, kv, o, wqb, wkvb, woutb = ctx.saved_tensors
        B, S, n_media, H, scale = ctx.dims
        M, D = y2.shape
        V, Dv = vis2.shape
        I = 64 * H
        dev = y2.device
        do = dout.to(torch.bfloat16).contiguous()
        d_o = torch.empty((M, I), dtype=torch.bfloat16, device=dev)
        _gemm(M, I, D, do, D, 0, woutb, I, 1, d_o)                                           # dout Wout
        dwout = torch.empty((D, I), dtype=torch.float32, device=dev)
        _gemm(D, I, M, do, D, 1, o, I, 1, dwout)                                             # dout^T o
        dq = torch.empty((M, I), dtype=torch.bfloat16, device=dev)
        dkv = torch.empty((V, 2 * I), dtype=torch.bfloat16, device=dev)
        check(_lib.load().fm_xattn_core_bwd(_ptr(q), _ptr(kv), _ptr(tt), _ptr(d_o), _ptr(dq), _ptr(dkv), B, S, n_media, H, float(scale),
                                            _stream()), "fm_xattn_core_bwd")
        dwq = torch.empty((I, D), dtype=torch.float32, device=dev)
        _gemm(I, D, M, dq, I, 1, yn, D, 1, dwq)                                              # dq^T yn
        dyn = torch.empty((M, D), dtype=torch.bfloat16, device=dev)
        _gemm(M, D, I, dq, I, 0, wqb, D, 1, dyn)                                             # dq Wq
        dwkv = torch.empty((2 * I, Dv), dtype=torch.float32, device=dev)
        _gemm(2 * I, Dv, V, dkv, 2 * I, 1, vis2, Dv, 1, dwkv)                                # dkv^T vis
        dvis = torch.empty((V, Dv), dtype=torch.bfloat16, device=dev)
        _gemm(V, Dv, 2 * I, dkv, 2 * I, 0, wkvb, Dv, 1, dvis)                                # dkv Wkv
        dy, dg, db = _ln_bwd(dyn, y2, g32, mean, rstd)
        t = ctx.dtypes
        return dy, dvis, None, dg.to(t[0]), db.to(t[1]), dwq.to(t[2]), dwkv.to(t[3]), dwout.to(t[4]), None, None


class _PerceiverAttentionFn(torch.autograd.Function):
    """PerceiverAttentionLayer (perceiver_resampler.py:32-96) with gradients: keys/values are [media ; latents], both normalised."""

    @staticmethod
    def forward(ctx, f2, l2, nm_w, nm_b, nl_w, nl_b, wq, wk, wv, wout, dims, norm_media, norm_latents):
        b, n1, H, scale = dims
        D = f2.shape[1]
        I = 64 * H
        dev = f2.device
        x, gm, mean_m, rstd_m = _ln_fwd_stats(f2, norm_media)
        lat, gl, mean_l, rstd_l = _ln_fwd_stats(l2, norm_latents)
        wqb, wkvb, woutb = _bf16(wq), _bf16(torch.cat([wk.detach(), wv.detach()], dim=0)), _bf16(wout)
        R, nk = b * 64, n1 + 64
        q = torch.empty((R, I), dtype=torch.bfloat16, device=dev)
        d = GemmDesc(M=R, N=I, K=D, A=_ptr(lat), lda=D, a_mn=0, B=_ptr(wqb), ldb=D, b_mn=0, epi=EPI_STORE, out=_ptr(q), ldo=I, out_f32=0, scale=scale)
        check(_lib.load().fm_gemm_bf16(C.byref(d), _stream()), "fm_gemm_bf16")
        kv_in = torch.cat([x.view(b, n1, D), lat.view(b, 64, D)], dim=1).reshape(b * nk, D)    # :65
        kv = torch.empty((b * nk, 2 * I), dtype=torch.bfloat16, device=dev)
        _gemm(b * nk, 2 * I, D, kv_in, D, 0, wkvb, D, 0, kv)
        o = torch.empty((R, I), dtype=torch.bfloat16, device=dev)
        lse = torch.empty(b * H * 64, dtype=torch.float32, device=dev)
        check(_lib.load().fm_resampler_core_fwd(_ptr(q), _ptr(kv), _ptr(o), _ptr(lse), b, nk, H, _stream()), "fm_resampler_core_fwd")
        out = torch.empty((R, D), dtype=l2.dtype, device=dev)
        _gemm(R, D, I, o, I, 0, woutb, I, 0, out)
        ctx.save_for_backward(f2, l2, gm, mean_m, rstd_m, gl, mean_l, rstd_l, lat, kv_in, q, kv, o, lse, wqb, wkvb, woutb)
        ctx.dims, ctx.dtypes = dims, (nm_w.dtype, nm_b.dtype, nl_w.dtype, nl_b.dtype, wq.dtype, wk.dtype, wv.dtype, wout.dtype)
        return out

    @staticmethod
    def backward(ctx, dout):
        f2, l2, gm, mean_m, rstd_m, gl, mean_l, rstd_l, lat, kv_in, q, kv, o, lse, wqb, wkvb, woutb = ctx.saved_tensors
        b, n1, H, scale = ctx.dims
        D = f2.shape[1]
        I = 64 * H
        R, nk = b * 64, n1 + 64
        dev = f2.device
        do = dout.to(torch.bfloat16).contiguous()
        d_o = torch.empty((R, I), dtype=torch.bfloat16, device=dev)
        _gemm(R, I, D, do, D, 0, woutb, I, 1, d_o)
        dwout = torch.empty((D, I), dtype=torch.float32, device=dev)
        _gemm(D, I, R, do, D, 1, o, I, 1, dwout)
        dq = torch.empty((R, I), dtype=torch.bfloat16, device=dev)
        dkv = torch.empty((b * nk, 2 * I), dtype=torch.bfloat16, device=dev)
        check(_lib.load().fm_resampler_core_bwd(_ptr(q), _ptr(kv), _ptr(o), _ptr(d_o), _ptr(lse), _ptr(dq), _ptr(dkv), b, nk, H,
                                                float(scale), _stream()), "fm_resampler_core_bwd")
        dwq = torch.empty((I, D), dtype=torch.float32, device=dev)
        _gemm(I, D, R, dq, I, 1, lat, D, 1, dwq)
        dlat_q = torch.empty((R, D), dtype=torch.bfloat16, device=dev)
        _gemm(R, D, I, dq, I, 0, wqb, D, 1, dlat_q)
        dwkv = torch.empty((2 * I, D), dtype=torch.float32, device=dev)
        _gemm(2 * I, D, b * nk, dkv, 2 * I, 1, kv_in, D, 1, dwkv)
        dkv_in = torch.empty((b * nk, D), dtype=torch.bfloat16, device=dev)
        _gemm(b * nk, D, 2 * I, dkv, 2 * I, 0, wkvb, D, 1, dkv_in)
        dkv_in = dkv_in.view(b, nk, D)
        dx = dkv_in[:, :n1].reshape(b * n1, D).contiguous()
        dlat = (dkv_in[:, n1:].reshape(R, D).float() + dlat_q.float()).to(torch.bfloat16).contiguous()   # latents feed q AND k/v
        dfeat, dgm, dbm = _ln_bwd(dx, f2, gm, mean_m, rstd_m)
        dl, dgl, dbl = _ln_bwd(dlat, l2, gl, mean_l, rstd_l)
        t = ctx.dtypes
        return (dfeat, dl, dgm.to(t[0]), dbm.to(t[1]), dgl.to(t[2]), dbl.to(t[3]), dwq.to(t[4]), dwkv[:I].to(t[5]), dwkv[I:].to(t[6]),
                dwout.to(t[7]), None, None, None)


def _wants_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def masked_cross_attention(mod, y: torch.Tensor, media_locations: torch.Tensor, visual_features: Optional[torch.Tensor],
                           previous_kv=None, output_kv: bool = False):
    """gated_cross_attention.py:42-131: returns (out (B,S,D), (k, v) | None) — the un-gated attention branch."""
    from .gated_cross_attention import _kv_buffer, _kv_views
    _need("fm_xattn_core_fwd")
    H = mod.heads
    if mod.n_visual != 64 or mod.to_q.weight.shape[0] != 64 * H:
        raise FlamingoB200Error("kernels are specialised for dim_head=64, n_visual=64")
    B, S, D = y.shape
    y2 = _as_rows(y, "MaskedCrossAttention input")
    if _wants_grad(y, visual_features, *mod.parameters()) and previous_kv is None and _lib.has("fm_xattn_core_bwd"):
        assert visual_features is not None and visual_features.ndim == 4                               # :84
        n_media = visual_features.shape[1]
        vis2 = visual_features.to(torch.bfloat16).reshape(-1, visual_features.shape[-1]).contiguous()
        tt = text_time_of(media_locations)
        out, kv = _MaskedCrossAttentionFn.apply(y2, vis2, tt, mod.norm.weight, mod.norm.bias, mod.to_q.weight, mod.to_kv.weight,
                                                mod.to_out.weight, (B, S, n_media, H, float(mod.scale)), mod.norm)
        return out.view(B, S, D), (_kv_views(kv, B, H, 64) if output_kv else None)
    _inference_only("MaskedCrossAttention", y, visual_features, *mod.parameters())
    q = _linear(_layernorm(y2, mod.norm), mod.to_q.weight, scale=mod.scale)                          # :74-78
    if previous_kv is None:
        assert visual_features is not None and visual_features.ndim == 4                               # :84
        vis = visual_features.to(torch.bfloat16).reshape(-1, visual_features.shape[-1]).contiguous()
        kv = _linear(vis, mod.to_kv.weight)                                                            # :86
    else:
        kv = _kv_buffer(*previous_kv)                                                                  # :90-92
    n_media = kv.shape[0] // (B * 64)
    tt = text_time_of(media_locations)                                                                 # :97
    if tt.shape[1] != S:                                                                               # cached decoding :102-104
        tt = tt[:, -S:].contiguous()
    o = torch.empty((B * S, 64 * H), dtype=torch.bfloat16, device=y.device)
    check(_lib.load().fm_xattn_core_fwd(_ptr(q), _ptr(kv), _ptr(tt), _ptr(o), B, S, n_media, H, _stream()), "fm_xattn_core_fwd")
    out = _linear(o, mod.to_out.weight, out_dtype=y.dtype).view(B, S, D)                               # :126
    return out, (_kv_views(kv, B, H, 64) if output_kv else None)


def perceiver_attention(mod, features: torch.Tensor, latents: torch.Tensor) -> torch.Tensor:
    """perceiver_resampler.py:32-96: features (b, n1, D), latents (b, 64, D) -> (b, 64, D)."""
    _need("fm_resampler_core_fwd")
    assert features.ndim == 3 and latents.ndim == 3 and features.shape[0] == latents.shape[0]        # :42-45
    assert features.shape[2] == latents.shape[2]
    b, n1, D = features.shape
    H = mod.heads
    if latents.shape[1] != 64 or mod.dim_head != 64:
        raise FlamingoB200Error("kernels are specialised for dim_head=64, 64 latents")
    if _wants_grad(features, latents, *mod.parameters()) and _lib.has("fm_resampler_core_bwd"):
        out = _PerceiverAttentionFn.apply(_as_rows(features, "features"), _as_rows(latents, "latents"), mod.norm_media.weight,
                                          mod.norm_media.bias, mod.norm_latents.weight, mod.norm_latents.bias, mod.to_q.weight,
                                          mod.to_k.weight, mod.to_v.weight, mod.to_out.weight, (b, n1, H, float(mod.scale)),
                                          mod.norm_media, mod.norm_latents)
        return out.view(b, 64, D)
    _inference_only("PerceiverAttentionLayer", features, latents, *mod.parameters())
    x = _layernorm(_as_rows(features, "features"), mod.norm_media).view(b, n1, D)                      # :52
    lat = _layernorm(_as_rows(latents, "latents"), mod.norm_latents)                                   # :53
    q = _linear(lat, mod.to_q.weight, scale=mod.scale)                                                 # :57, :79
    kv_in = torch.cat([x, lat.view(b, 64, D)], dim=1).reshape(b * (n1 + 64), D)                        # :65
    w_kv = torch.cat([mod.to_k.weight.detach(), mod.to_v.weight.detach()], dim=0)                      # :69-70 as one operand
    kv = _linear(kv_in, w_kv)
    o = torch.empty((b * 64, 64 * H), dtype=torch.bfloat16, device=features.device)
    check(_lib.load().fm_resampler_core_fwd(_ptr(q), _ptr(kv), _ptr(o), None, b, n1 + 64, H, _stream()), "fm_resampler_core_fwd")
    return _linear(o, mod.to_out.weight, out_dtype=latents.dtype).view(b, 64, D)                       # :96
