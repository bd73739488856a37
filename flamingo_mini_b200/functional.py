"""torch.autograd bridges onto the C ABI (fm_resampler_{fwd,bwd}, fm_xattn_{fwd,bwd}).

Host code is PyTorch only for memory, streams and autograd bookkeeping: every tensor (outputs, activations kept
for backward, scratch, gradients) is allocated here with torch's caching allocator and handed to the library as a
raw device pointer together with the current CUDA stream.  The library never allocates and keeps no references.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import ACT_IDS, FlamingoB200Error, ResamplerCfg, ResamplerLayout, XattnCfg, XattnLayout, check


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise FlamingoB200Error(
            f"{what}: tensor is on {t.device}; flamingo_mini_b200 only runs on an sm_100a CUDA device "
            "(there is no CPU or eager fallback)")


class FlatParams:
    """One contiguous fp32 buffer holding every parameter of a module in the library's layout; the module's
    nn.Parameters are views into it, so the bf16 tensor-core shadow is one cast kernel and the gradient arena one
    tensor (what the data-parallel reducer all-reduces)."""

    def __init__(self, total: int, slots: Sequence[tuple]):
        # slots: (parameter, offset) in registration order
        self.total = int(total)
        self.slots = list(slots)
        self.flat: Optional[torch.Tensor] = None
        self._shadow: Optional[torch.Tensor] = None
        self._shadow_ver = None
        self.always_refresh = False      # True: re-cast on every forward (what a step after an optimizer update does)

    def params(self):
        return [p for p, _ in self.slots]

    def attach(self) -> None:
        """(Re)build the flat buffer on the parameters' current device and re-point the parameters into it."""
        ps = self.params()
        dev = ps[0].device
        flat = torch.zeros(self.total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, off in self.slots:
                n = p.numel()
                flat[off:off + n].copy_(p.detach().reshape(-1).to(torch.float32))
                p.data = flat[off:off + n].view(p.shape)
        self.flat = flat
        self._shadow = None
        self._shadow_ver = None

    def is_attached(self) -> bool:
        if self.flat is None:
            return False
        base = self.flat.data_ptr()
        return all(p.dtype == torch.float32 and p.data_ptr() == base + 4 * off for p, off in self.slots)

    def ensure(self) -> torch.Tensor:
        if not self.is_attached():
            self.attach()
        return self.flat

    def shadow_bf16(self) -> torch.Tensor:
        """bf16 copy of the flat buffer, refreshed when any parameter was modified in place."""
        flat = self.ensure()
        ver = sum(p._version for p, _ in self.slots)
        if self.always_refresh or self._shadow is None or self._shadow_ver != ver or self._shadow.device != flat.device:
            if self._shadow is None or self._shadow.device != flat.device:
                self._shadow = torch.empty(self.total, dtype=torch.bfloat16, device=flat.device)
            _require_cuda(flat, "parameter cast")
            check(_lib.load().fm_cast_f32_to_bf16(_ptr(flat), _ptr(self._shadow), self.total, _stream()), "fm_cast_f32_to_bf16")
            self._shadow_ver = ver
        return self._shadow

    def invalidate_shadow(self) -> None:
        """The flat buffer was updated in place (optimizer step, checkpoint load): re-cast on the next forward."""
        self._shadow_ver = None

    def grad_views(self, g: torch.Tensor):
        return [g[off:off + p.numel()].view(p.shape) for p, off in self.slots]

    def grad_arena(self, owner) -> torch.Tensor:
        """The module's flat fp32 gradient arena (same layout as `flat`).  It is allocated ONCE per module and reused by every
        backward while no parameter still holds a `.grad` (the usual `zero_grad(set_to_none=True)` training loop; a CUDA-graph
        capture then sees a static address).  The layout's total is rounded up to 8 floats and the kernels never write that
        tail, so the arena starts zeroed: norms / all-reduces / optimizer moments over the whole arena see zeros there, never
        stale allocator bytes.  When gradients are being accumulated into existing `.grad`s (which alias the cached arena), a
        fresh zeroed arena is returned instead."""
        flat = self.ensure()
        cached = getattr(owner, "_grad_arena", None)
        free = all(p.grad is None for p in self.params())
        if cached is not None and free and cached.device == flat.device and cached.numel() == self.total:
            return cached
        g = torch.zeros(self.total, dtype=torch.float32, device=flat.device)
        if free:
            owner._grad_arena = g
        return g

    def current_grad(self, owner):
        """The gradient as it stands in the parameters' `.grad` right now, in the arena layout: (flat fp32 tensor, aliased).
        Normally every `.grad` is a view of the module's cached arena and that arena itself is returned (aliased=True: writing
        into it IS writing the `.grad`s).  After gradient accumulation the cached arena holds the accumulated sum - not the arena
        the LAST backward wrote, which only held that micro-batch's part.  If the `.grad`s were assigned from elsewhere, a
        gathered copy is returned (aliased=False; `scatter_grad` writes it back).  (None, True) when no parameter has a gradient."""
        ps = self.params()
        if all(p.grad is None for p in ps):
            return None, True
        for cand in (getattr(owner, "_grad_arena", None), getattr(owner, "_last_grad_arena", None)):
            if cand is not None and cand.numel() == self.total and all(
                    p.grad is not None and p.grad.dtype == torch.float32 and p.grad.is_contiguous()
                    and p.grad.data_ptr() == cand.data_ptr() + 4 * off for p, off in self.slots):
                return cand, True
        flat = self.ensure()
        g = torch.zeros(self.total, dtype=torch.float32, device=flat.device)
        for p, off in self.slots:
            if p.grad is not None:
                g[off:off + p.numel()].copy_(p.grad.detach().reshape(-1))
        return g, False

    def scatter_grad(self, g: torch.Tensor) -> None:
        """Write a flat gradient (arena layout) back into the parameters' `.grad` (counterpart of a non-aliased current_grad)."""
        with torch.no_grad():
            for p, off in self.slots:
                if p.grad is not None:
                    p.grad.copy_(g[off:off + p.numel()].view(p.shape))


# ============================================================================================ gated xattn block
def xattn_cfg(B, S, D, Dv, n_media, heads, dim_head, ff_inner, act, y_f32, training) -> XattnCfg:
    return XattnCfg(B=B, S=S, D=D, Dv=Dv, n_media=n_media, heads=heads, dim_head=dim_head, ff_inner=ff_inner,
                    act=ACT_IDS[act], y_f32=int(y_f32), training=int(training))


def xattn_layout(D, Dv, heads, dim_head, ff_inner) -> XattnLayout:
    cfg = xattn_cfg(1, 1, D, Dv, 1, heads, dim_head, ff_inner, "gelu", 0, 0)
    L = XattnLayout()
    check(_lib.load().fm_xattn_layout_of(cfg, L), "fm_xattn_layout_of")
    return L


def text_time_of(media_locations: torch.Tensor) -> torch.Tensor:
    """int32 running count of <image> markers (gated_cross_attention.py:97) computed by fm_text_time."""
    _require_cuda(media_locations, "media_locations")
    ml = media_locations.to(torch.int32).contiguous()
    tt = torch.empty_like(ml)
    check(_lib.load().fm_text_time(_ptr(ml), _ptr(tt), ml.shape[0], ml.shape[1], _stream()), "fm_text_time")
    return tt


# ---- deferred side-stream join (fm_set_option("defer_join", 1) / FM_B200_OPTS=defer_join=1) --------------------
# With the switch on, fm_xattn_bwd returns while its weight-gradient GEMMs are still running on the library's side stream, so
# they overlap the frozen LM block's backward that autograd runs next.  Until side_join() every buffer those kernels touch must
# stay allocated and the parameter gradients must not be read: the backward below parks the buffers here, and whoever consumes
# gradients (GradArenaReducer, training.train, bench.py; a CUDA-graph capture cannot even end without it) calls side_join().
_PENDING: list = []


def defer_join_enabled() -> bool:
    """Asks the library itself (the switch may have been set through FM_B200_OPTS or fm_set_option directly)."""
    return _lib.has("fm_get_option") and _lib.load().fm_get_option(_lib.OPTION_KEYS["defer_join"]) == 1


def set_defer_join(on: bool) -> bool:
    """Turns the deferred join on or off; returns False when the loaded library does not export it (stale build)."""
    if not _lib.has("fm_side_join") or _lib.load().fm_set_option(_lib.OPTION_KEYS["defer_join"], int(bool(on))) != 0:
        return False
    if not on:
        side_join()
    return True


def side_join() -> None:
    """The current stream waits for the library's side stream; the buffers parked by deferred backwards are released."""
    if _lib.has("fm_side_join"):
        check(_lib.load().fm_side_join(_stream()), "fm_side_join")
    _PENDING.clear()


class _XattnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, y, vis, text_time, kv_given, *params):
        lib = _lib.load()
        fp: FlatParams = mod._fp
        B, S, D = y.shape
        training = any(ctx.needs_input_grad)      # grad mode itself is off inside Function.forward
        y_f32 = y.dtype == torch.float32
        if kv_given is None:
            n_media = vis.shape[1]
            vis2 = vis.reshape(B * n_media * vis.shape[2], vis.shape[3])
            kv = torch.empty((B * n_media * 64, 2 * mod.heads * mod.dim_head), dtype=torch.bfloat16, device=y.device)
        else:
            kv = kv_given
            n_media = kv.shape[0] // (B * 64)
            vis2 = None
        cfg = xattn_cfg(B, S, D, mod.dim_visual, n_media, mod.heads, mod.dim_head, mod.ff_inner, mod.act, y_f32, training)
        w_f32 = fp.ensure()
        w_bf16 = fp.shadow_bf16()
        saved = torch.empty(lib.fm_xattn_saved_bytes(cfg), dtype=torch.uint8, device=y.device)
        y_out = torch.empty_like(y)
        check(lib.fm_xattn_fwd(cfg, _ptr(w_f32), _ptr(w_bf16), _ptr(y), _ptr(vis2), _ptr(text_time), _ptr(kv),
                               int(kv_given is not None), _ptr(y_out), _ptr(saved), _stream()), "fm_xattn_fwd")
        if training:
            ctx.mod, ctx.cfg = mod, cfg
            ctx.vis_shape = None if vis is None else vis.shape
            ctx.had_vis = vis2 is not None
            ctx.save_for_backward(y, vis2 if vis2 is not None else y.new_empty(0), text_time, kv, saved, w_bf16)
        ctx.mark_non_differentiable(kv)
        return y_out, kv

    @staticmethod
    def backward(ctx, dy_out, _dkv):
        lib = _lib.load()
        mod, cfg = ctx.mod, ctx.cfg
        y, vis2, text_time, kv, saved, w_bf16 = ctx.saved_tensors
        fp: FlatParams = mod._fp
        dy_out = dy_out.contiguous()
        dy = torch.empty_like(y)
        dvis = torch.empty_like(vis2) if ctx.had_vis else None
        g = fp.grad_arena(mod)
        scratch = torch.empty(lib.fm_xattn_scratch_bytes(cfg), dtype=torch.uint8, device=y.device)
        check(lib.fm_xattn_bwd(cfg, _ptr(fp.flat), _ptr(w_bf16), _ptr(y), _ptr(vis2) if ctx.had_vis else None,
                               _ptr(text_time), _ptr(kv), _ptr(saved), _ptr(dy_out), _ptr(dy), _ptr(dvis), _ptr(g),
                               _ptr(scratch), _stream()), "fm_xattn_bwd")
        mod._last_grad_arena = g
        if defer_join_enabled():
            _PENDING.append((dy_out, saved, scratch, vis2, kv, y, text_time, w_bf16, g))
            if any(p.grad is not None for p in fp.params()):      # autograd would ADD into .grad right away: join first
                side_join()
        hook = getattr(mod, "_grad_ready_hook", None)
        if hook is not None:
            hook(mod, g)
        dvis_full = dvis.view(ctx.vis_shape) if dvis is not None else None
        return (None, dy, dvis_full, None, None, *fp.grad_views(g))


def xattn_block(mod, y: torch.Tensor, visual_features: Optional[torch.Tensor], text_time: torch.Tensor,
                kv: Optional[torch.Tensor]):
    """y: (B,S,D) bf16/fp32; visual_features: (B,N,64,Dv) or None when kv is given; returns (y_out, kv[B*N*64, 2*heads*dim_head])."""
    _require_cuda(y, "GatedCrossAttentionBlock input")
    if y.dtype not in (torch.bfloat16, torch.float32):
        raise FlamingoB200Error(f"unsupported activation dtype {y.dtype}: use bfloat16 or float32")
    y = y.contiguous()
    if kv is None:
        vis = visual_features.to(torch.bfloat16).contiguous()
    else:
        vis = None
    # Attach BEFORE apply(): autograd records each parameter's dtype when apply() collects its inputs.  A module that was cast
    # (`lm.to(torch.bfloat16)` reaches the blocks inside the LM layers) would otherwise be recorded as bf16, re-pointed at its
    # fp32 master copy inside forward, and the engine would convert every gradient to bf16 copies on that first step.
    mod._fp.ensure()
    return _XattnFn.apply(mod, y, vis, text_time, kv, *mod._fp.params())


# ============================================================================================ perceiver resampler
def resampler_cfg(BN, T, F, Dv, depth, heads, dim_head, n_latents, n_time_embeds, ff_inner, act, x_f32, training):
    return ResamplerCfg(BN=BN, T=T, F=F, Dv=Dv, depth=depth, heads=heads, dim_head=dim_head, n_latents=n_latents,
                        n_time_embeds=n_time_embeds, ff_inner=ff_inner, act=ACT_IDS[act], x_f32=int(x_f32),
                        training=int(training))


def resampler_layout(Dv, depth, heads, dim_head, n_latents, n_time_embeds, ff_inner) -> ResamplerLayout:
    cfg = resampler_cfg(1, 1, 1, Dv, depth, heads, dim_head, n_latents, n_time_embeds, ff_inner, "gelu", 0, 0)
    L = ResamplerLayout()
    check(_lib.load().fm_resampler_layout_of(cfg, L), "fm_resampler_layout_of")
    return L


class _ResamplerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x_f, out_dtype, *params):
        lib = _lib.load()
        fp: FlatParams = mod._fp
        BN, T, F, Dv = x_f.shape
        training = any(ctx.needs_input_grad)
        cfg = resampler_cfg(BN, T, F, Dv, mod.depth, mod.heads, mod.dim_head, mod.n_queries, mod.num_time_embeds,
                            mod.ff_inner, mod.act, x_f.dtype == torch.float32, training)
        w_f32 = fp.ensure()
        w_bf16 = fp.shadow_bf16()
        nbytes = lib.fm_resampler_saved_bytes(cfg)
        if nbytes == 0:
            check(lib.fm_resampler_layout_of(cfg, ResamplerLayout()), "resampler configuration")
        saved = torch.empty(nbytes, dtype=torch.uint8, device=x_f.device)
        out = torch.empty((BN, mod.n_queries, Dv), dtype=out_dtype, device=x_f.device)
        check(lib.fm_resampler_fwd(cfg, _ptr(w_f32), _ptr(w_bf16), _ptr(x_f), _ptr(out), int(out_dtype == torch.float32),
                                   _ptr(saved), _stream()), "fm_resampler_fwd")
        if training:
            ctx.mod, ctx.cfg = mod, cfg
            ctx.save_for_backward(x_f, saved, w_bf16)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        mod, cfg = ctx.mod, ctx.cfg
        x_f, saved, w_bf16 = ctx.saved_tensors
        fp: FlatParams = mod._fp
        dout = dout.to(torch.bfloat16).contiguous()
        g = fp.grad_arena(mod)
        scratch = torch.empty(lib.fm_resampler_scratch_bytes(cfg), dtype=torch.uint8, device=x_f.device)
        layer_hook = getattr(mod, "_grad_layer_hook", None)
        hook = getattr(mod, "_grad_ready_hook", None)
        mod._last_grad_arena = g
        if layer_hook is not None and _lib.has("fm_resampler_bwd_notify"):
            # per-layer completion (fm_resampler_bwd_notify): each layer's slice of the arena is handed over while backward still runs
            failure = []

            def _layer_done(_user, layer):
                try:
                    lo, hi = mod._layer_ranges[layer]
                    layer_hook(mod, g, lo, hi)
                except BaseException as e:          # an exception must not unwind through the C frame
                    failure.append(e)

            cb = _lib.LAYER_CB(_layer_done)
            check(lib.fm_resampler_bwd_notify(cfg, _ptr(fp.flat), _ptr(w_bf16), _ptr(x_f), _ptr(saved), _ptr(dout), _ptr(g),
                                              _ptr(scratch), cb, None, _stream()), "fm_resampler_bwd_notify")
            if failure:
                raise failure[0]
            if hook is not None:                     # what is left: latents / time_pos_emb before the layers, final norm after
                hook(mod, g, [(0, mod._layer_ranges[0][0]), (mod._layer_ranges[-1][1], fp.total)])
        else:
            check(lib.fm_resampler_bwd(cfg, _ptr(fp.flat), _ptr(w_bf16), _ptr(x_f), _ptr(saved), _ptr(dout), _ptr(g),
                                       _ptr(scratch), _stream()), "fm_resampler_bwd")
            if hook is not None:
                hook(mod, g)
        return (None, None, None, *fp.grad_views(g))


def resampler(mod, x_f: torch.Tensor) -> torch.Tensor:
    """x_f: (BN, T, F, Dv) bf16/fp32 -> (BN, 64, Dv) in x_f's dtype."""
    _require_cuda(x_f, "PerceiverResampler input")
    if x_f.dtype not in (torch.bfloat16, torch.float32):
        raise FlamingoB200Error(f"unsupported feature dtype {x_f.dtype}: use bfloat16 or float32")
    if x_f.requires_grad:
        raise FlamingoB200Error("PerceiverResampler: gradient w.r.t. the CLIP features is not produced "
                                "(they are computed under no_grad in the reference, modeling_flamingo.py:169-170); "
                                "detach x_f")
    mod._fp.ensure()          # before apply(): see xattn_block
    return _ResamplerFn.apply(mod, x_f.contiguous(), x_f.dtype, *mod._fp.params())


# ============================================================================================ loss head
class _CrossEntropyFn(torch.autograd.Function):
    """mean over counted rows of (lse - logit[target]) through fm_cross_entropy_{fwd,bwd} (modeling_flamingo.py:287-298)."""

    @staticmethod
    def forward(ctx, logits, targets, vocab, ignore_index):
        lib = _lib.load()
        rows, ld = logits.shape
        lse = torch.empty(rows, dtype=torch.float32, device=logits.device)
        row_loss = torch.empty(rows, dtype=torch.float32, device=logits.device)
        check(lib.fm_cross_entropy_fwd(_ptr(logits), ld, rows, vocab, _ptr(targets), ignore_index, _ptr(lse), _ptr(row_loss),
                                       _stream()), "fm_cross_entropy_fwd")
        count = (targets != ignore_index).sum().clamp_(min=1).to(torch.float32)      # stays on the device: graph-capturable
        ctx.save_for_backward(logits, targets, lse, count)
        ctx.vocab, ctx.ignore_index = vocab, ignore_index
        return (row_loss.sum() / count).to(logits.dtype)

    @staticmethod
    def backward(ctx, dloss):
        lib = _lib.load()
        logits, targets, lse, count = ctx.saved_tensors
        rows, ld = logits.shape
        scale = (dloss.to(torch.float32) / count).reshape(1).contiguous()
        dlogits = torch.empty_like(logits)
        check(lib.fm_cross_entropy_bwd(_ptr(logits), ld, rows, ctx.vocab, _ptr(targets), ctx.ignore_index, _ptr(lse),
                                       _ptr(scale), _ptr(dlogits), _stream()), "fm_cross_entropy_bwd")
        return dlogits, None, None, None


def cross_entropy(logits: torch.Tensor, targets: torch.Tensor, vocab: int, ignore_index: int = -100) -> torch.Tensor:
    """logits: (rows, ld) bf16 with ld % 8 == 0 (columns >= vocab are padding); targets: (rows,) int64. Mean reduction."""
    _require_cuda(logits, "cross_entropy logits")
    if logits.dtype != torch.bfloat16 or logits.ndim != 2 or not logits.is_contiguous():
        raise FlamingoB200Error("cross_entropy: logits must be a contiguous 2-D bfloat16 tensor")
    return _CrossEntropyFn.apply(logits, targets.to(torch.int64).contiguous(), int(vocab), int(ignore_index))
