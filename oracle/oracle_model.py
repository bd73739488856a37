"""CPU restatement of the reference's MODEL-LEVEL training step (TEST / BASELINE INFRASTRUCTURE ONLY).

`bench.py --impl reference` and the `cpu_baseline` leg time this: the oracle restatements of PerceiverResampler and
GatedCrossAttentionBlock (oracle/flamingo_oracle.py, pinned against the unmodified reference by tests/golden) spliced into the
stock HuggingFace language model exactly the way the reference splices its own modules.  Nothing here imports
`flamingo_mini_b200`: the reference arm must not map the product's shared library.

Reference lines followed (flamingo_mini/…):
  modeling_flamingo.py:76-94    every xattn_every-th LM layer is wrapped (`_init_layers`)
  modeling_flamingo.py:105-119  the LM is frozen except its input embedding (tied with lm_head) and the xattn blocks
  modeling_flamingo.py:323/348  resize_token_embeddings(vocab + 1) for <EOC>
  modeling_flamingo.py:241-279  condition() every wrapped block, run the LM, lm_head
  modeling_flamingo.py:287-298  shifted next-token cross entropy
  gated_cross_attention.py:214-252  ModifiedLMBlock.condition / forward (with the *args pass-through transformers >= 5 needs for
                                GPT-2 blocks; body unchanged)
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from .flamingo_oracle import OracleGatedXattn, OracleResampler


class OracleModifiedLMBlock(nn.Module):
    """gated_cross_attention.py:187-252"""

    def __init__(self, lm_block, **kw):
        super().__init__()
        self.xattn_block = OracleGatedXattn(**kw)
        self.lm_block = lm_block
        self.visual_features = None
        self.media_locations = None
        self.xattn_layer_past = None
        self.kv_output = None

    def condition(self, visual_features, media_locations, xattn_layer_past=None):
        self.visual_features, self.media_locations, self.xattn_layer_past = visual_features, media_locations, xattn_layer_past

    def forward(self, hidden_states, *args, use_cache=False, **kwargs):
        hidden_states, kv = self.xattn_block(hidden_states, self.visual_features, self.media_locations,
                                             previous_kv=self.xattn_layer_past, output_kv=bool(use_cache))
        self.kv_output = kv
        return self.lm_block(hidden_states, *args, use_cache=use_cache, **kwargs)


class OracleFlamingo(nn.Module):
    """resampler + HF LM with gated xattn blocks + lm_head + loss, for `visual_features`-driven training steps (the CLIP tower is
    bypassed by the benchmark: inputs are CLIP patch features, modeling_flamingo.py:189,241)."""

    def __init__(self, lm: str, lm_config: dict, dim: int, dim_visual: int, xattn_every: int = 1, resampler_depth: int = 6,
                 fused_gelu: bool = True, seed: int = 0, alpha: float = 0.5):
        super().__init__()
        torch.manual_seed(seed)
        if lm.startswith("gpt"):
            from transformers import GPT2Config, GPT2LMHeadModel
            base = GPT2LMHeadModel(GPT2Config(**lm_config))
            assert dim == base.config.n_embd
            base.resize_token_embeddings(base.config.vocab_size + 1)
            if fused_gelu:          # same formula as HF's NewGELUActivation in one op (the B200 arm's frozen LM does the same)
                from transformers.activations import NewGELUActivation
                for module in base.modules():
                    for name, child in list(module.named_children()):
                        if isinstance(child, NewGELUActivation):
                            setattr(module, name, nn.GELU(approximate="tanh"))
            self.lm, self.lm_head, layers = base.transformer, base.lm_head, base.transformer.h
        else:
            from transformers import OPTConfig, OPTForCausalLM
            base = OPTForCausalLM(OPTConfig(**lm_config))
            assert dim == base.config.hidden_size
            base.resize_token_embeddings(base.config.vocab_size + 1)
            self.lm, self.lm_head, layers = base.model, base.lm_head, base.model.decoder.layers
        self.resampler = OracleResampler(dim=dim_visual, depth=resampler_depth, seed=seed)
        self.modified = []
        for i, idx in enumerate(range(0, len(layers), xattn_every)):
            layers[idx] = OracleModifiedLMBlock(layers[idx], dim=dim, dim_visual=dim_visual, seed=seed + 1 + i)
            self.modified.append(layers[idx])
        for p in self.lm.parameters():                       # freeze_lm (modeling_flamingo.py:105-119)
            p.requires_grad = False
        self.lm.get_input_embeddings().weight.requires_grad = True
        for layer in self.modified:
            for p in layer.xattn_block.parameters():
                p.requires_grad = True
            with torch.no_grad():                            # at the reference's init (alpha = 0) every block is the identity
                layer.xattn_block.alpha_attn.fill_(alpha)
                layer.xattn_block.alpha_ffw.fill_(alpha)

    def training_step(self, clip_feats, input_ids, media_locations, n_images: int):
        """fwd + bwd of one batch; returns the loss.  clip_feats (B*N, T, F, Dv)."""
        B = input_ids.shape[0]
        vf = self.resampler(clip_feats)
        vf = vf.reshape(B, n_images, vf.shape[-2], vf.shape[-1])
        for layer in self.modified:
            layer.condition(vf, media_locations, None)
        out = self.lm(input_ids=input_ids, attention_mask=torch.ones_like(input_ids), use_cache=False, return_dict=True)
        logits = self.lm_head(out.last_hidden_state)
        loss = F.cross_entropy(logits[..., :-1, :].reshape(-1, logits.size(-1)), input_ids[..., 1:].reshape(-1))
        loss.backward()
        return loss
