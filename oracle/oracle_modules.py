"""Swap the CPU oracle into a FlamingoModel (TEST / BASELINE INFRASTRUCTURE ONLY).

Used by ``bench.py --impl reference`` / the ``cpu_baseline`` leg and by CPU plumbing tests: replaces the CUDA
PerceiverResampler and every ModifiedLMBlock.xattn_block by the nn.Module faces of the oracle
(oracle/flamingo_oracle.py), keeping the model-level code path identical.  Never imported by flamingo_mini_b200.
"""
from __future__ import annotations

from .flamingo_oracle import OracleGatedXattn, OracleResampler


def _copy_named(src_module, oracle_module):
    """Copy parameters of a (reference-named) module into the oracle module ('a.b.c' -> attribute 'a__b__c')."""
    import torch
    with torch.no_grad():
        for name, p in src_module.named_parameters():
            getattr(oracle_module, name.replace(".", "__")).copy_(p)


def swap_in_oracle(model, seed: int = 0, copy_weights: bool = False):
    """copy_weights=True keeps the weights of the modules being replaced (used by the model-level golden test)."""
    fl = model.flamingo
    c = fl.config
    old_resampler = fl.resampler
    fl.resampler = OracleResampler(dim=c.dim_visual, depth=c.resampler_depth, dim_head=c.resampler_dim_head,
                                   heads=c.resampler_heads, num_latents=c.resampler_num_latents,
                                   num_time_embeds=c.resampler_num_time_embeds, ff_mult=c.resampler_ff_mult,
                                   act=c.resampler_act, seed=seed)
    if copy_weights:
        _copy_named(old_resampler, fl.resampler)
    for i, layer in enumerate(fl.get_modified_layers()):
        old_block = layer.xattn_block
        layer.xattn_block = OracleGatedXattn(dim=c.dim, dim_visual=c.dim_visual, dim_head=c.xattn_dim_head,
                                             heads=c.xattn_heads, ff_mult=c.xattn_ff_mult, act=c.xattn_act,
                                             n_visual=c.resampler_num_latents, seed=seed + 1 + i)
        if copy_weights:
            _copy_named(old_block, layer.xattn_block)
    return model
