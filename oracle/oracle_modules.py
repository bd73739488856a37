"""Swap the CPU oracle into a FlamingoModel (TEST / BASELINE INFRASTRUCTURE ONLY).

Used by ``bench.py --impl reference`` / the ``cpu_baseline`` leg and by CPU plumbing tests: replaces the CUDA
PerceiverResampler and every ModifiedLMBlock.xattn_block by the nn.Module faces of the oracle
(oracle/flamingo_oracle.py), keeping the model-level code path identical.  Never imported by flamingo_mini_b200.
"""
from __future__ import annotations

from .flamingo_oracle import OracleGatedXattn, OracleResampler


def swap_in_oracle(model, seed: int = 0):
    fl = model.flamingo
    c = fl.config
    fl.resampler = OracleResampler(dim=c.dim_visual, depth=c.resampler_depth, dim_head=c.resampler_dim_head,
                                   heads=c.resampler_heads, num_latents=c.resampler_num_latents,
                                   num_time_embeds=c.resampler_num_time_embeds, ff_mult=c.resampler_ff_mult,
                                   act=c.resampler_act, seed=seed)
    for i, layer in enumerate(fl.get_modified_layers()):
        layer.xattn_block = OracleGatedXattn(dim=c.dim, dim_visual=c.dim_visual, dim_head=c.xattn_dim_head,
                                             heads=c.xattn_heads, ff_mult=c.xattn_ff_mult, act=c.xattn_act,
                                             n_visual=c.resampler_num_latents, seed=seed + 1 + i)
    return model
