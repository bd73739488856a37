"""Golden vectors for the MODEL-LEVEL host glue (modeling_flamingo.py of the reference): run the unmodified reference
FlamingoModel (OPT branch; the GPT-2 branch raises TypeError under transformers >= 5) with tiny random-init
HF models (from_pretrained patched to build from configs - there is no hub access) and record logits / loss / cache shapes.

    python tests/golden/make_golden_model.py        # build container only (needs /root/reference)
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("FLAMINGO_REF", "/root/reference")

OPT_CFG = dict(hidden_size=64, num_hidden_layers=2, num_attention_heads=2, ffn_dim=128, vocab_size=97,
               max_position_embeddings=64, word_embed_proj_dim=64, dropout=0.0, attention_dropout=0.0, activation_dropout=0.0)
CLIP_CFG = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=1, num_attention_heads=2, image_size=32, patch_size=16,
                attention_dropout=0.0)


def main():
    from einops import rearrange, repeat
    shim = types.ModuleType("einops_exts")
    shim.rearrange_many = lambda ts, pat, **kw: tuple(rearrange(t, pat, **kw) for t in ts)
    shim.repeat_many = lambda ts, pat, **kw: tuple(repeat(t, pat, **kw) for t in ts)
    sys.modules["einops_exts"] = shim
    sys.path.insert(0, REF)
    import transformers
    from transformers import CLIPVisionConfig, CLIPVisionModel, OPTConfig, OPTForCausalLM
    CLIPVisionModel.from_pretrained = classmethod(lambda cls, *a, **k: cls(CLIPVisionConfig(**CLIP_CFG)))
    OPTForCausalLM.from_pretrained = classmethod(lambda cls, *a, **k: cls(OPTConfig(**OPT_CFG)))
    from flamingo_mini import FlamingoConfig, FlamingoModel        # the reference package

    torch.manual_seed(0)
    cfg = FlamingoConfig(lm="facebook/opt-125m", dim=64, dim_visual=64, xattn_every=1, resampler_depth=1)
    model = FlamingoModel(cfg).eval()
    with torch.no_grad():
        for layer in model.flamingo.get_modified_layers():
            layer.xattn_block.alpha_attn.fill_(0.4)
            layer.xattn_block.alpha_ffw.fill_(-0.3)
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(0, 97, (2, 11), generator=g)
    ml = torch.zeros(2, 11, dtype=torch.long)
    ml[0, 2] = 1
    ml[1, 0] = 1
    ml[1, 6] = 1                                            # second marker, only one image -> uniform rows
    pix = torch.randn(2, 1, 3, 32, 32, generator=g)
    out = model(input_ids=ids, media_locations=ml, pixel_values=pix, labels=ids, attention_mask=torch.ones_like(ids))
    out.loss.backward()
    trainable = sorted(model.state_dict_trainable().keys())
    fx = dict(opt_cfg=OPT_CFG, clip_cfg=CLIP_CFG, state_dict={k: v.clone() for k, v in model.state_dict().items()},
              input_ids=ids, media_locations=ml, pixel_values=pix, logits=out.logits.detach(), loss=out.loss.detach(),
              trainable_keys=trainable, n_trainable=sum(p.numel() for p in model.parameters_trainable()),
              grad_alpha_attn_layer0=model.flamingo.lm.decoder.layers[0].xattn_block.alpha_attn.grad.clone(),
              grad_latents=model.flamingo.resampler.latents.grad.clone(), transformers=transformers.__version__)
    # every trainable gradient (resampler, gated xattn blocks, token embedding): the GPU model-level parity test compares all of them
    fx["grads"] = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.requires_grad and p.grad is not None}
    # cached forward: prefix then one more token through (xattn, lm) caches
    with torch.no_grad():
        first = model(input_ids=ids[:, :8], media_locations=ml[:, :8], pixel_values=pix, use_cache=True,
                      attention_mask=torch.ones_like(ids[:, :8]))
        fx["cache_k_shape"] = tuple(first.past_key_values[0][0][0].shape)
        fx["logits_prefix"] = first.logits
        # one cached decode step (modeling_flamingo.py:238-239,264-265; gated_cross_attention.py:88-104): token 8 with both caches
        step = model(input_ids=ids[:, 8:9], media_locations=ml[:, :9], past_key_values=first.past_key_values, use_cache=True,
                     attention_mask=torch.ones_like(ids[:, :9]))
        fx["logits_step"] = step.logits
    torch.save(fx, os.path.join(HERE, "model_opt_tiny.pt"))
    print("saved; logits", tuple(out.logits.shape), "loss", float(out.loss.detach()), "trainable", fx["n_trainable"], len(trainable),
          "cache k", fx["cache_k_shape"], "size", os.path.getsize(os.path.join(HERE, "model_opt_tiny.pt")))

    # ---- GPT-2 branch (the benchmark's LM family).  transformers >= 5 calls GPT-2 blocks with positional extras, which
    # the reference's ModifiedLMBlock.forward(hidden_states, use_cache=False, **kwargs) cannot accept (TypeError).
    # Only that signature is widened here (same body, same arithmetic); everything else is the unmodified reference.
    from transformers import GPT2Config, GPT2LMHeadModel
    from flamingo_mini import gated_cross_attention as ref_gx

    def fwd(self, hidden_states, *args, use_cache=False, **kwargs):
        hidden_states, kv = self.xattn_block(y=hidden_states, visual_features=self.visual_features,
                                             media_locations=self.media_locations, previous_kv=self.xattn_layer_past,
                                             output_kv=use_cache)
        self.kv_output = kv
        return self.lm_block(hidden_states, *args, use_cache=use_cache, **kwargs)
    ref_gx.ModifiedLMBlock.forward = fwd
    GPT2_CFG = dict(n_embd=64, n_layer=2, n_head=2, vocab_size=97, n_positions=64, resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)
    GPT2LMHeadModel.from_pretrained = classmethod(lambda cls, *a, **k: cls(GPT2Config(**GPT2_CFG)))
    torch.manual_seed(2)
    cfg = FlamingoConfig(lm="gpt2", dim=64, dim_visual=64, xattn_every=2, resampler_depth=1, xattn_act="sqrelu")
    model = FlamingoModel(cfg).eval()
    with torch.no_grad():
        for layer in model.flamingo.get_modified_layers():
            layer.xattn_block.alpha_attn.fill_(-0.6)
            layer.xattn_block.alpha_ffw.fill_(0.2)
    out = model(input_ids=ids, media_locations=ml, pixel_values=pix, labels=ids, attention_mask=torch.ones_like(ids))
    out.loss.backward()
    fx2 = dict(gpt2_cfg=GPT2_CFG, clip_cfg=CLIP_CFG, state_dict={k: v.clone() for k, v in model.state_dict().items()},
               input_ids=ids, media_locations=ml, pixel_values=pix, logits=out.logits.detach(), loss=out.loss.detach(),
               trainable_keys=sorted(model.state_dict_trainable().keys()),
               n_modified=len(list(model.flamingo.get_modified_layers())),
               grad_alpha_ffw_layer0=model.flamingo.lm.h[0].xattn_block.alpha_ffw.grad.clone(),
               grads={n: p.grad.detach().clone() for n, p in model.named_parameters() if p.requires_grad and p.grad is not None},
               note="reference ModifiedLMBlock.forward signature widened with *args for transformers>=5 (body unchanged)")
    with torch.no_grad():
        first = model(input_ids=ids[:, :8], media_locations=ml[:, :8], pixel_values=pix, use_cache=True,
                      attention_mask=torch.ones_like(ids[:, :8]))
        fx2["logits_prefix"] = first.logits
        step = model(input_ids=ids[:, 8:9], media_locations=ml[:, :9], past_key_values=first.past_key_values, use_cache=True,
                     attention_mask=torch.ones_like(ids[:, :9]))
        fx2["logits_step"] = step.logits
    torch.save(fx2, os.path.join(HERE, "model_gpt2_tiny.pt"))
    print("saved gpt2; logits", tuple(out.logits.shape), "loss", float(out.loss.detach()), "modified layers", fx2["n_modified"])


if __name__ == "__main__":
    main()
