"""CPU oracle for the flamingo-mini hot path (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

This file is a from-scratch *functional* restatement of the arithmetic of the two
trainable modules of dhansmair/flamingo-mini.  It exists only to check the CUDA
kernels: nothing in ``flamingo_mini_b200/`` may import it.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` use it.

Parity status: the reference ships no golden vectors or numerical tests
("parity unpinned by the reference").  The oracle is therefore pinned
against *outputs of the reference itself*, run in the build container:
``tests/golden/make_golden.py`` imports the unmodified reference modules from
``/root/reference`` (with a two-function ``einops_exts`` shim), records their
forward outputs and autograd gradients on seeded inputs into ``tests/golden/*.pt``
and ``tests/test_oracle_golden.py`` checks this file against those fixtures.

All functions take a flat ``params`` dict whose keys are the reference's state-dict
names (relative to the module), so a reference checkpoint can be fed directly.
Any floating dtype works; parity tests run it in fp64/fp32.

Reference line citations are relative to ``/root/reference/flamingo_mini/``.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

LN_EPS = 1e-5  # torch.nn.LayerNorm default used at perceiver_resampler.py:24-25,140
                # gated_cross_attention.py:36, utils.py:46


# ----------------------------------------------------------------------------- helpers
def _ln(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """LayerNorm over the last axis, biased variance, eps inside the sqrt."""
    mu = x.mean(dim=-1, keepdim=True)
    xc = x - mu
    var = (xc * xc).mean(dim=-1, keepdim=True)
    return xc * torch.rsqrt(var + LN_EPS) * w + b


def _act(x: torch.Tensor, act: str) -> torch.Tensor:
    """utils.py:36-40 — gelu is the exact erf form (nn.GELU default), sqrelu = relu(x)^2."""
    if act == "gelu":
        return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))
    if act == "sqrelu":
        r = torch.clamp_min(x, 0.0)
        return r * r
    if act == "relu":
        return torch.clamp_min(x, 0.0)
    raise AssertionError(f"act. can only be one of gelu/sqrelu/relu, got {act}")


def feed_forward(x: torch.Tensor, p: Params, prefix: str, act: str = "gelu") -> torch.Tensor:
    """utils.py:31-50: Sequential(LayerNorm, Linear(no bias), act, Linear(no bias)).

    Parameter names: ``{prefix}0.weight|bias`` (LN), ``{prefix}1.weight`` (inner,dim),
    ``{prefix}3.weight`` (dim,inner).
    """
    h = _ln(x, p[prefix + "0.weight"], p[prefix + "0.bias"])
    h = h @ p[prefix + "1.weight"].transpose(0, 1)
    h = _act(h, act)
    return h @ p[prefix + "3.weight"].transpose(0, 1)


def _split_heads(t: torch.Tensor, heads: int) -> torch.Tensor:
    """(b, n, h*d) -> (b, h, n, d)"""
    b, n, hd = t.shape
    return t.reshape(b, n, heads, hd // heads).permute(0, 2, 1, 3)


def _merge_heads(t: torch.Tensor) -> torch.Tensor:
    """(b, h, n, d) -> (b, n, h*d)"""
    b, h, n, d = t.shape
    return t.permute(0, 2, 1, 3).reshape(b, n, h * d)


def _stable_softmax(sim: torch.Tensor) -> torch.Tensor:
    """Row max is subtracted as a constant (detached), perceiver_resampler.py:88-89,
    gated_cross_attention.py:114-115."""
    sim = sim - sim.amax(dim=-1, keepdim=True).detach()
    return torch.softmax(sim, dim=-1)


# ----------------------------------------------------------------------------- resampler
def perceiver_attention(features: torch.Tensor, latents: torch.Tensor, p: Params, prefix: str,
                        heads: int = 8, dim_head: int = 64) -> torch.Tensor:
    """perceiver_resampler.py:32-96.

    features (b, f, d) and latents (b, q, d).  Keys/values are computed from
    ``[LN_media(features) ; LN_latents(latents)]`` (concatenated on the sequence axis,
    :65), queries from ``LN_latents(latents)`` only (:57); the query is scaled by
    dim_head**-0.5 *before* the QK^T product (:79).  No mask, no dropout, no biases.
    """
    assert features.ndim == 3 and latents.ndim == 3
    assert features.shape[0] == latents.shape[0] and features.shape[2] == latents.shape[2]
    xm = _ln(features, p[prefix + "norm_media.weight"], p[prefix + "norm_media.bias"])
    xl = _ln(latents, p[prefix + "norm_latents.weight"], p[prefix + "norm_latents.bias"])
    q = xl @ p[prefix + "to_q.weight"].transpose(0, 1)
    kv_in = torch.cat((xm, xl), dim=-2)
    k = kv_in @ p[prefix + "to_k.weight"].transpose(0, 1)
    v = kv_in @ p[prefix + "to_v.weight"].transpose(0, 1)
    q, k, v = _split_heads(q, heads), _split_heads(k, heads), _split_heads(v, heads)
    q = q * (dim_head ** -0.5)
    alphas = _stable_softmax(q @ k.transpose(-1, -2))
    out = _merge_heads(alphas @ v)
    return out @ p[prefix + "to_out.weight"].transpose(0, 1)


def perceiver_resampler(x_f: torch.Tensor, p: Params, depth: int, heads: int = 8, dim_head: int = 64,
                        act: str = "gelu", prefix: str = "") -> torch.Tensor:
    """perceiver_resampler.py:143-188.

    x_f: (b, n, d) or (b, T, n, d).  Adds ``time_pos_emb[:T]`` (shape (num_time_embeds,1,d),
    :166 — T > num_time_embeds is a broadcasting error), flattens (T n) (:172), repeats the
    learned latents over the batch (:179), runs depth x {x += attn(x_f, x); x += ffw(x)}
    (:181-183) and a final LayerNorm (:187).  Returns (b, num_latents, d).
    """
    if x_f.ndim == 3:
        x_f = x_f.unsqueeze(1)
    assert x_f.ndim == 4
    b, T, n, d = x_f.shape
    latents = p[prefix + "latents"]
    assert d == latents.shape[1]
    tpe = p[prefix + "time_pos_emb"]
    if T > tpe.shape[0]:
        raise RuntimeError(f"n_frames={T} exceeds num_time_embeds={tpe.shape[0]} (perceiver_resampler.py:166)")
    x_f = (x_f + tpe[:T]).reshape(b, T * n, d)
    x = latents.unsqueeze(0).expand(b, -1, -1)
    for i in range(depth):
        x = x + perceiver_attention(x_f, x, p, f"{prefix}layers.{i}.0.", heads, dim_head)
        x = x + feed_forward(x, p, f"{prefix}layers.{i}.1.", act)
    assert x.shape == (b, latents.shape[0], d)
    return _ln(x, p[prefix + "norm.weight"], p[prefix + "norm.bias"])


# ----------------------------------------------------------------------------- gated xattn
def text_time_of(media_locations: torch.Tensor) -> torch.Tensor:
    """gated_cross_attention.py:97 — running count of <image> markers per row."""
    return media_locations.cumsum(dim=-1)


def masked_cross_attention(y: torch.Tensor, media_locations: torch.Tensor,
                           visual_features: Optional[torch.Tensor], p: Params, prefix: str,
                           heads: int = 8, dim_head: int = 64, n_visual: int = 64,
                           previous_kv: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                           output_kv: bool = False):
    """gated_cross_attention.py:42-131.

    Token i may attend only to the n_visual latents of image number ``text_time[i]``
    (1-based, equality at :111).  Masked scores are filled with -finfo.max (:112).
    Rows with ``text_time == 0`` have their attention zeroed *after* the softmax
    (:119-121).  Rows whose text_time exceeds the number of images are fully masked,
    so after the row-max subtraction they become a uniform average over all keys.
    ``to_kv`` weight rows [0, inner) produce K and [inner, 2*inner) produce V (:86).
    """
    _, n_token, _ = y.shape
    yn = _ln(y, p[prefix + "norm.weight"], p[prefix + "norm.bias"])
    q = (yn @ p[prefix + "to_q.weight"].transpose(0, 1)) * (dim_head ** -0.5)
    q = _split_heads(q, heads)
    if previous_kv is None:
        bv, n_media = visual_features.shape[:2]
        vis = visual_features.reshape(bv, n_media * visual_features.shape[2], visual_features.shape[3])
        kvp = vis @ p[prefix + "to_kv.weight"].transpose(0, 1)
        inner = kvp.shape[-1] // 2
        k, v = _split_heads(kvp[..., :inner], heads), _split_heads(kvp[..., inner:], heads)
    else:
        k, v = previous_kv
        n_media = k.shape[2] // n_visual
    sim = q @ k.transpose(-1, -2)                                   # (b, h, i, j)
    text_time = text_time_of(media_locations)
    if previous_kv is not None:
        text_time = text_time[:, -n_token:]                          # :102-104
        assert text_time.shape == y.shape[:2]
    media_time = torch.arange(n_media, device=y.device).repeat_interleave(n_visual) + 1
    allowed = text_time[:, None, :, None] == media_time[None, None, None, :]
    sim = sim.masked_fill(~allowed, -torch.finfo(sim.dtype).max)
    alphas = _stable_softmax(sim)
    alphas = alphas.masked_fill((text_time == 0)[:, None, :, None], 0.0)
    out = _merge_heads(alphas @ v) @ p[prefix + "to_out.weight"].transpose(0, 1)
    return (out, (k, v)) if output_kv else (out, None)


def gated_xattn_block(y: torch.Tensor, visual_features: Optional[torch.Tensor], media_locations: torch.Tensor,
                      p: Params, prefix: str = "", heads: int = 8, dim_head: int = 64, n_visual: int = 64,
                      act: str = "gelu", previous_kv=None, output_kv: bool = False):
    """gated_cross_attention.py:160-184: y += tanh(alpha_attn)*attn(y); y += tanh(alpha_ffw)*ffw(y)."""
    if previous_kv is None:
        assert visual_features.ndim == 4
    attn_out, kv = masked_cross_attention(y, media_locations, visual_features, p, prefix + "attn.",
                                          heads, dim_head, n_visual, previous_kv, output_kv)
    y = y + torch.tanh(p[prefix + "alpha_attn"]) * attn_out
    y = y + torch.tanh(p[prefix + "alpha_ffw"]) * feed_forward(y, p, prefix + "ffw.", act)
    return y, kv


# ----------------------------------------------------------------------------- parameter factories
def resampler_param_shapes(dim: int, depth: int, dim_head: int = 64, heads: int = 8, num_latents: int = 64,
                           num_time_embeds: int = 4, ff_mult: int = 4):
    """Name -> shape, in the reference's registration order (perceiver_resampler.py:128-140)."""
    inner, ffi = dim_head * heads, int(dim * ff_mult)
    shapes = {"latents": (num_latents, dim), "time_pos_emb": (num_time_embeds, 1, dim)}
    for i in range(depth):
        a, f = f"layers.{i}.0.", f"layers.{i}.1."
        shapes.update({
            a + "norm_media.weight": (dim,), a + "norm_media.bias": (dim,),
            a + "norm_latents.weight": (dim,), a + "norm_latents.bias": (dim,),
            a + "to_q.weight": (inner, dim), a + "to_k.weight": (inner, dim),
            a + "to_v.weight": (inner, dim), a + "to_out.weight": (dim, inner),
            f + "0.weight": (dim,), f + "0.bias": (dim,),
            f + "1.weight": (ffi, dim), f + "3.weight": (dim, ffi),
        })
    shapes.update({"norm.weight": (dim,), "norm.bias": (dim,)})
    return shapes


def xattn_param_shapes(dim: int, dim_visual: int, dim_head: int = 64, heads: int = 8, ff_mult: int = 4):
    """Name -> shape, reference order (gated_cross_attention.py:36-40,154-158)."""
    inner, ffi = dim_head * heads, int(dim * ff_mult)
    return {
        "attn.norm.weight": (dim,), "attn.norm.bias": (dim,),
        "attn.to_q.weight": (inner, dim), "attn.to_kv.weight": (2 * inner, dim_visual),
        "attn.to_out.weight": (dim, inner),
        "alpha_attn": (1,),
        "ffw.0.weight": (dim,), "ffw.0.bias": (dim,),
        "ffw.1.weight": (ffi, dim), "ffw.3.weight": (dim, ffi),
        "alpha_ffw": (1,),
    }


def seeded_params(shapes: Dict[str, tuple], seed: int, dtype=torch.float32, alpha: float = 0.5) -> Params:
    """Deterministic test parameters (NOT the reference initialiser): matrices ~ N(0,1)/sqrt(fan_in),
    LN weights ~ 1 + 0.1 N(0,1), LN/other biases ~ 0.1 N(0,1), latents/time embeddings ~ N(0,1),
    gates = ``alpha`` (at the reference's init of 0 every non-gate gradient vanishes).
    Generated in fp64 from a private CPU generator so fixtures are reproducible."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    out: Params = {}
    for name, shape in shapes.items():
        if name.startswith("alpha"):
            t = torch.full(shape, alpha, dtype=torch.float64)
        elif len(shape) == 1:
            r = torch.randn(shape, generator=g, dtype=torch.float64) * 0.1
            t = r + 1.0 if name.endswith("weight") else r
        elif name in ("latents", "time_pos_emb"):
            t = torch.randn(shape, generator=g, dtype=torch.float64)
        else:
            t = torch.randn(shape, generator=g, dtype=torch.float64) / math.sqrt(shape[-1])
        out[name] = t.to(dtype)
    return out


# ----------------------------------------------------------------------------- nn.Module adapters
class OracleResampler(torch.nn.Module):
    """nn.Module face of :func:`perceiver_resampler` with the reference's parameter names;
    used only as the CPU baseline / checker."""

    def __init__(self, *, dim, depth, dim_head=64, heads=8, num_latents=64, num_time_embeds=4, ff_mult=4,
                 act="gelu", seed: int = 0):
        super().__init__()
        self.depth, self.heads, self.dim_head, self.act = depth, heads, dim_head, act
        shapes = resampler_param_shapes(dim, depth, dim_head, heads, num_latents, num_time_embeds, ff_mult)
        self._names = list(shapes)
        for n, t in seeded_params(shapes, seed).items():
            self.register_parameter(n.replace(".", "__"), torch.nn.Parameter(t))

    def params(self) -> Params:
        return {n: getattr(self, n.replace(".", "__")) for n in self._names}

    def forward(self, x_f):
        return perceiver_resampler(x_f, self.params(), self.depth, self.heads, self.dim_head, self.act)


class OracleGatedXattn(torch.nn.Module):
    """nn.Module face of :func:`gated_xattn_block` (checker / CPU baseline only)."""

    def __init__(self, *, dim, dim_visual, dim_head=64, heads=8, ff_mult=4, act="gelu", n_visual=64, seed: int = 0):
        super().__init__()
        self.heads, self.dim_head, self.act, self.n_visual = heads, dim_head, act, n_visual
        shapes = xattn_param_shapes(dim, dim_visual, dim_head, heads, ff_mult)
        self._names = list(shapes)
        for n, t in seeded_params(shapes, seed).items():
            self.register_parameter(n.replace(".", "__"), torch.nn.Parameter(t))

    def params(self) -> Params:
        return {n: getattr(self, n.replace(".", "__")) for n in self._names}

    def forward(self, y, visual_features, media_locations, previous_kv=None, output_kv=False):
        return gated_xattn_block(y, visual_features, media_locations, self.params(), "", self.heads,
                                 self.dim_head, self.n_visual, self.act, previous_kv, output_kv)
