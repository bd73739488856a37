"""GatedCrossAttentionBlock / MaskedCrossAttention / ModifiedLMBlock with the reference's module API
(flamingo_mini/gated_cross_attention.py), computed by the sm_100a kernels behind ``fm_xattn_fwd`` / ``fm_xattn_bwd``.

Parameter names match the reference (``alpha_attn``, ``alpha_ffw``, ``attn.norm.*``, ``attn.to_q.weight``,
``attn.to_kv.weight``, ``attn.to_out.weight``, ``ffw.{0,1,3}.*``).  ``ModifiedLMBlock.forward`` additionally accepts
the positional arguments transformers >= 5 passes to GPT-2 blocks, and forwards them untouched.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import nn

from . import functional as Fn
from .utils import FeedForward


class MaskedCrossAttention(nn.Module):
    """The attention half of the block (gated_cross_attention.py:15-40): parameters + a stand-alone inference forward."""

    def __init__(self, *, dim, dim_visual, dim_head=64, heads=8, n_visual=64):
        super().__init__()
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.n_visual = n_visual
        inner_dim = dim_head * heads
        self.norm = nn.LayerNorm(dim)
        self.to_q = nn.Linear(dim, inner_dim, bias=False)
        self.to_kv = nn.Linear(dim_visual, inner_dim * 2, bias=False)
        self.to_out = nn.Linear(inner_dim, dim, bias=False)

    def forward(self, y, media_locations, visual_features, previous_kv=None, output_kv=False):
        """Stand-alone inference forward (gated_cross_attention.py:42-131; note the argument order differs from the
        block's).  Training goes through GatedCrossAttentionBlock, where this runs fused with the gates and the FFW."""
        from .standalone import masked_cross_attention
        return masked_cross_attention(self, y, media_locations, visual_features, previous_kv, output_kv)


def _kv_views(kv: torch.Tensor, B: int, heads: int, dim_head: int):
    """[B*V, 2*H*dh] buffer -> reference-shaped (k, v), each (B, H, V, dh) (gated_cross_attention.py:86-87)."""
    V = kv.shape[0] // B
    t = kv.view(B, V, 2, heads, dim_head)
    return t[:, :, 0].permute(0, 2, 1, 3), t[:, :, 1].permute(0, 2, 1, 3)


def _kv_buffer(k: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """Reference-shaped cached (k, v) -> the library's [B*V, 2*H*dh] bf16 buffer (no copy when they are our own views)."""
    B, H, V, dh = k.shape
    W = 2 * H * dh
    if (k.dtype == torch.bfloat16 and v.dtype == torch.bfloat16 and k.stride() == (V * W, dh, W, 1)
            and v.stride() == k.stride() and v.data_ptr() == k.data_ptr() + 2 * H * dh):
        return k.as_strided((B * V, W), (W, 1), k.storage_offset())
    return torch.cat([k.permute(0, 2, 1, 3).reshape(B, V, H * dh), v.permute(0, 2, 1, 3).reshape(B, V, H * dh)],
                     dim=-1).reshape(B * V, W).to(torch.bfloat16).contiguous()


class GatedCrossAttentionBlock(nn.Module):
    def __init__(self, *, dim, dim_visual, dim_head=64, heads=8, ff_mult=4, act="gelu", n_visual=64):
        super().__init__()
        self.attn = MaskedCrossAttention(dim=dim, dim_visual=dim_visual, dim_head=dim_head, heads=heads, n_visual=n_visual)
        self.alpha_attn = nn.Parameter(torch.tensor([0.]))
        self.ffw = FeedForward(dim, mult=ff_mult, act=act)
        self.alpha_ffw = nn.Parameter(torch.tensor([0.]))

        self.dim, self.dim_visual, self.dim_head, self.heads = dim, dim_visual, dim_head, heads
        self.n_visual, self.act, self.ff_inner = n_visual, act, int(dim * ff_mult)
        if n_visual != 64:
            raise ValueError("flamingo_mini_b200 kernels are specialised for n_visual = 64 latents per image")
        L = Fn.xattn_layout(dim, dim_visual, heads, dim_head, self.ff_inner)
        self._fp = Fn.FlatParams(L.total, [
            (self.attn.norm.weight, L.attn_norm_w), (self.attn.norm.bias, L.attn_norm_b),
            (self.attn.to_q.weight, L.to_q), (self.attn.to_kv.weight, L.to_kv), (self.attn.to_out.weight, L.to_out),
            (self.alpha_attn, L.alpha_attn),
            (self.ffw[0].weight, L.ffw_norm_w), (self.ffw[0].bias, L.ffw_norm_b),
            (self.ffw[1].weight, L.ffw_w1), (self.ffw[3].weight, L.ffw_w2),
            (self.alpha_ffw, L.alpha_ffw),
        ])
        self._grad_ready_hook = None
        self._last_grad_arena = None

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self._fp.flat = None
        return out

    def forward(self, y: torch.Tensor, visual_features: Optional[torch.Tensor], media_locations: torch.Tensor,
                previous_kv: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, output_kv: bool = False,
                text_time: Optional[torch.Tensor] = None):
        """(gated_cross_attention.py:160-184)
        y (n_batch, n_tokens, d_token); visual_features (n_batch, n_media, n_queries, dim_visual);
        media_locations (n_batch, n_tokens) bool/int.  Returns (y, (k, v) | None).
        text_time (extension, optional): int32 cumsum(media_locations) (:97) already computed by the caller — it depends only on
        media_locations, which every block of a forward pass shares, so FlamingoBaseModel.forward computes it once and hands
        it to the blocks through ModifiedLMBlock.condition()."""
        if previous_kv is None:
            assert visual_features is not None and visual_features.ndim == 4
        shape_before = y.shape
        tt = text_time if text_time is not None else Fn.text_time_of(media_locations)
        kv_in = None
        if previous_kv is not None:
            kv_in = _kv_buffer(*previous_kv)
            n_token = y.shape[1]
            if tt.shape[1] != n_token:                      # cached decoding: last n_token positions (:102-104)
                tt = tt[:, -n_token:].contiguous()
                assert tt.shape == y.shape[:2]
        y_out, kv = Fn.xattn_block(self, y, visual_features, tt, kv_in)
        assert y_out.shape == shape_before
        return y_out, (_kv_views(kv, y.shape[0], self.heads, self.dim_head) if output_kv else None)


class ModifiedLMBlock(nn.Module):
    """Gated cross-attention followed by the wrapped LM block (gated_cross_attention.py:187-252)."""

    def __init__(self, lm_block, **kwargs):
        super().__init__()
        self.xattn_block = GatedCrossAttentionBlock(**kwargs)
        self.lm_block = lm_block
        self.visual_features = None
        self.media_locations = None
        self.xattn_layer_past = None
        self.text_time = None
        self.kv_output = None

    def condition(self, visual_features: torch.Tensor, media_locations: torch.Tensor, xattn_layer_past=None, text_time=None) -> None:
        """Side channel set by the model before the LM runs (gated_cross_attention.py:214-229).  `text_time` (optional
        extension): the shared int32 cumsum of media_locations, see GatedCrossAttentionBlock.forward."""
        self.visual_features = visual_features
        self.media_locations = media_locations
        self.xattn_layer_past = xattn_layer_past
        self.text_time = text_time

    def forward(self, hidden_states, *args, use_cache: Optional[bool] = False, **kwargs):
        hidden_states, kv = self.xattn_block(
            y=hidden_states,
            visual_features=self.visual_features,
            media_locations=self.media_locations,
            previous_kv=self.xattn_layer_past,
            output_kv=bool(use_cache),
            **({} if self.text_time is None else {"text_time": self.text_time}),
        )
        self.kv_output = kv
        return self.lm_block(hidden_states, *args, use_cache=use_cache, **kwargs)
