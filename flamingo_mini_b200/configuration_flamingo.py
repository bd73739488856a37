"""FlamingoConfig — same fields and defaults as flamingo_mini/configuration_flamingo.py:6-27 of the reference
(the 18 constructor keywords are the checkpoint/config.json contract), plus two optional offline hooks."""
from __future__ import annotations

from transformers.configuration_utils import PretrainedConfig

_DEFAULTS = dict(
    lm="gpt2",
    clip_model_type="openai/clip-vit-base-patch32",
    dim=1024,
    dim_visual=768,
    xattn_every=1,
    xattn_dim_head=64,
    xattn_heads=8,
    xattn_ff_mult=4,
    xattn_act="gelu",
    resampler_depth=6,
    resampler_dim_head=64,
    resampler_heads=8,
    resampler_num_latents=64,
    resampler_num_time_embeds=4,
    resampler_ff_mult=4,
    resampler_act="gelu",
    freeze_language_model=True,
    freeze_vision_model=True,
)


class FlamingoConfig(PretrainedConfig):
    """Configuration of a Flamingo model.

    lm / clip_model_type are HuggingFace identifiers ('gpt2*' or 'facebook/opt-*'; a CLIP vision tower);
    dim / dim_visual are the LM and vision widths; xattn_* configure the gated cross-attention blocks inserted
    before every ``xattn_every``-th LM layer; resampler_* configure the PerceiverResampler.

    Offline extension (not in the reference): ``lm_config`` / ``clip_config`` may hold HF config dicts; when given,
    the language model / vision tower are built from them with random weights instead of ``from_pretrained``
    (there is no hub access on the benchmark machines).  ``lm_fused_gelu`` (default True): HuggingFace evaluates GPT-2's
    ``gelu_new`` as eight separate elementwise kernels (pow, mul, add, tanh ...); the identical formula is available
    as ONE torch kernel, ``F.gelu(approximate="tanh")`` (difference 9e-16 in fp64), which the frozen LM then uses.
    ``fused_cross_entropy`` (default True): on CUDA bf16 logits with mean reduction the shifted next-token loss is computed
    by the library's row kernels (one pass forward keeping only the row log-sum-exp, one pass backward) instead of torch's
    log_softmax + nll_loss (measured -0.44 ms per C2 step on a B200, profiles/r02_validate_next); every other case (fp32
    logits, CPU, reduction != "mean") takes torch's path.
    """
    model_type = "flamingo"

    def __init__(self, lm_config: dict | None = None, clip_config: dict | None = None, lm_fused_gelu: bool = True,
                 fused_cross_entropy: bool = True, **kwargs):
        for name, default in _DEFAULTS.items():
            setattr(self, name, kwargs.pop(name, default))
        self.lm_config = lm_config
        self.clip_config = clip_config
        self.lm_fused_gelu = lm_fused_gelu
        # loss head through fm_cross_entropy_{fwd,bwd} where applicable (CUDA, bf16 logits, mean reduction)
        self.fused_cross_entropy = fused_cross_entropy
        super().__init__(**kwargs)
