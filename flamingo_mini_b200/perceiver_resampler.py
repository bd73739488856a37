"""PerceiverResampler with the reference's module API (flamingo_mini/perceiver_resampler.py:99-188), computed by
the sm_100a kernels behind ``fm_resampler_fwd`` / ``fm_resampler_bwd``.

Constructor keywords, ``forward(x_f)`` and every parameter name/shape match the reference, so its checkpoints load
unchanged: ``latents``, ``time_pos_emb``, ``layers.{i}.0.{norm_media,norm_latents,to_q,to_k,to_v,to_out}.*``,
``layers.{i}.1.{0,1,3}.*``, ``norm.*``.
"""
from __future__ import annotations

import torch
from torch import nn

from . import functional as Fn
from .utils import FeedForward


class PerceiverAttentionLayer(nn.Module):
    """One resampler attention layer (perceiver_resampler.py:9-30): parameters + a stand-alone inference forward."""

    def __init__(self, *, dim, dim_head=64, heads=8):
        super().__init__()
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.dim_head = dim_head
        inner_dim = dim_head * heads
        self.norm_media = nn.LayerNorm(dim)
        self.norm_latents = nn.LayerNorm(dim)
        self.to_q = nn.Linear(dim, inner_dim, bias=False)
        self.to_k = nn.Linear(dim, inner_dim, bias=False)
        self.to_v = nn.Linear(dim, inner_dim, bias=False)
        self.to_out = nn.Linear(inner_dim, dim, bias=False)

    def forward(self, features, latents):
        """Stand-alone inference forward (perceiver_resampler.py:32-96): features (b, n1, D), latents (b, 64, D).
        Training goes through PerceiverResampler, where the whole layer stack runs fused."""
        from .standalone import perceiver_attention
        return perceiver_attention(self, features, latents)


class PerceiverResampler(nn.Module):
    def __init__(self, *, dim, depth, dim_head=64, heads=8, num_latents=64, num_time_embeds=4, ff_mult=4, act="gelu"):
        super().__init__()
        self.dim = dim
        self.n_queries = num_latents
        self.depth, self.heads, self.dim_head = depth, heads, dim_head
        self.num_time_embeds, self.act = num_time_embeds, act
        self.ff_inner = int(dim * ff_mult)

        self.latents = nn.Parameter(torch.randn(num_latents, dim))
        self.time_pos_emb = nn.Parameter(torch.randn(num_time_embeds, 1, dim))
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(nn.ModuleList([
                PerceiverAttentionLayer(dim=dim, dim_head=dim_head, heads=heads),
                FeedForward(dim=dim, mult=ff_mult, act=act),
            ]))
        self.norm = nn.LayerNorm(dim)

        L = Fn.resampler_layout(dim, depth, heads, dim_head, num_latents, num_time_embeds, self.ff_inner)
        slots = [(self.latents, L.latents), (self.time_pos_emb, L.time_pos_emb)]
        for i, (attn, ffw) in enumerate(self.layers):
            b = L.layer0 + i * L.layer_stride
            slots += [(attn.norm_media.weight, b + L.norm_media_w), (attn.norm_media.bias, b + L.norm_media_b),
                      (attn.norm_latents.weight, b + L.norm_latents_w), (attn.norm_latents.bias, b + L.norm_latents_b),
                      (attn.to_q.weight, b + L.to_q), (attn.to_k.weight, b + L.to_k), (attn.to_v.weight, b + L.to_v),
                      (attn.to_out.weight, b + L.to_out),
                      (ffw[0].weight, b + L.ffw_norm_w), (ffw[0].bias, b + L.ffw_norm_b),
                      (ffw[1].weight, b + L.ffw_w1), (ffw[3].weight, b + L.ffw_w2)]
        slots += [(self.norm.weight, L.norm_w), (self.norm.bias, L.norm_b)]
        self._fp = Fn.FlatParams(L.total, slots)
        self._layer_ranges = [(int(L.layer0 + i * L.layer_stride), int(L.layer0 + (i + 1) * L.layer_stride)) for i in range(depth)]
        self._grad_ready_hook = None
        self._grad_layer_hook = None      # optional (module, arena, lo, hi): called per finished layer during backward
        self._last_grad_arena = None

    def _apply(self, fn, *a, **kw):           # .to()/.cuda() move parameters one by one: re-flatten lazily
        out = super()._apply(fn, *a, **kw)
        self._fp.flat = None
        return out

    def forward(self, x_f: torch.Tensor) -> torch.Tensor:
        """x_f: (b, n, d) or (b, T, n, d) CLIP features -> (b, num_latents, d) (perceiver_resampler.py:143-188)."""
        if x_f.ndim == 3:
            x_f = x_f.unsqueeze(1)
        assert x_f.ndim == 4
        assert x_f.shape[3] == self.dim
        if x_f.shape[1] > self.num_time_embeds:
            raise RuntimeError(f"n_frames={x_f.shape[1]} exceeds num_time_embeds={self.num_time_embeds} "
                               "(time_pos_emb broadcast, perceiver_resampler.py:166)")
        out = Fn.resampler(self, x_f)
        assert out.shape == torch.Size([x_f.shape[0], self.n_queries, self.dim])
        return out
