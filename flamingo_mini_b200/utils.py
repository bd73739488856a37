"""Helpers mirroring flamingo_mini/utils.py (reference): FeedForward / SquaredReLU containers and small I/O utils.

``FeedForward(dim, mult, act)`` keeps the reference's ``nn.Sequential(LayerNorm, Linear, act, Linear)`` structure
so parameter names (``0.weight``, ``0.bias``, ``1.weight``, ``3.weight``) and checkpoints are unchanged
(utils.py:31-50).  Inside PerceiverResampler / GatedCrossAttentionBlock these containers only *hold* parameters:
the arithmetic runs in the fused sm_100a kernels; called on their own they run a forward (and, when something requires
grad, a backward) composed from the library's LayerNorm / GEMM primitives (standalone.py).
"""
from __future__ import annotations

import torch
from torch import nn

ACTS = ("gelu", "sqrelu", "relu")


def load_url(url: str):
    import requests
    from PIL import Image
    return Image.open(requests.get(url, stream=True).raw)


def load_image(path: str):
    from PIL import Image
    return Image.open(path)


def unzip(l):
    return list(zip(*l))


class SquaredReLU(nn.Module):
    """relu(x)**2 (utils.py:22-28)."""

    def forward(self, x):
        return torch.relu(x).square()


class _FeedForward(nn.Sequential):
    """The reference's Sequential (same indices / parameter names); arithmetic in the sm_100a library."""

    def __init__(self, dim: int, mult: int = 4, act: str = "gelu"):
        assert act in ACTS, f"act. can only be one of {ACTS}"
        inner = int(dim * mult)
        act_mod = {"gelu": nn.GELU, "sqrelu": SquaredReLU, "relu": nn.ReLU}[act]()
        super().__init__(nn.LayerNorm(dim), nn.Linear(dim, inner, bias=False), act_mod, nn.Linear(inner, dim, bias=False))
        self.dim, self.inner_dim, self.act = dim, inner, act

    def forward(self, x):
        """Stand-alone forward (LayerNorm + tcgen05 GEMM with fused activation + GEMM), differentiable (standalone.py);
        inside the two hot-path modules the same arithmetic runs fused."""
        from .standalone import feed_forward
        return feed_forward(self, x)


def FeedForward(dim, mult=4, act="gelu"):
    return _FeedForward(dim, mult, act)


def get_common_prefix_length(x: torch.Tensor) -> int:
    """Number of leading columns on which all rows of the matrix x agree (utils.py:53-58)."""
    same = (x[:1] == x[1:]).all(dim=0)
    diff = (~same).nonzero()
    return int(diff[0]) if diff.numel() else x.size(1)
