"""FlamingoModel / FlamingoGPT2 / FlamingoOPT with the reference's public API
(flamingo_mini/modeling_flamingo.py) around the sm_100a PerceiverResampler and gated cross-attention blocks.

Host code is plain PyTorch + HuggingFace: the frozen CLIP encoder and language model are the stock HF modules
(north_star: "the frozen CLIP encoder and LM remain reference PyTorch"); only the resampler and the
``ModifiedLMBlock``s that are spliced into every ``xattn_every``-th LM layer run the new kernels.

API kept from the reference: ``FlamingoModel(config)``, ``forward(input_ids, attention_mask, media_locations,
pixel_values, visual_features, head_mask, inputs_embeds, use_cache, past_key_values, return_dict, labels,
loss_reduction)``, ``parameters_trainable``, ``state_dict_trainable``, ``freeze_lm/unfreeze_lm/freeze_vm``,
``prepare_inputs_for_generation``, ``_reorder_cache``, ``generate_captions``, ``score_sequences`` and the
``flamingo.{vision_encoder,resampler,lm,lm_head}`` sub-module names (checkpoint keys).
"""
from __future__ import annotations

import contextlib
import logging
from typing import Any, Dict, Iterable, List, Optional

import torch
import torch.nn.functional as F
from torch import nn
from transformers import PreTrainedModel
from transformers.modeling_outputs import CausalLMOutputWithPast

from .configuration_flamingo import FlamingoConfig
from .gated_cross_attention import ModifiedLMBlock
from .perceiver_resampler import PerceiverResampler
from .utils import get_common_prefix_length

try:  # transformers >= 4.50 no longer mixes generation into PreTrainedModel
    from transformers.generation import GenerationMixin
except Exception:  # pragma: no cover
    GenerationMixin = object


@contextlib.contextmanager
def suppress_model_loading_warnings(suppress: bool = True):
    log = logging.getLogger("transformers.modeling_utils")
    old = log.level
    if suppress:
        log.setLevel(logging.CRITICAL)
    try:
        yield
    finally:
        log.setLevel(old)


def _build_clip(config: FlamingoConfig):
    from transformers import CLIPVisionConfig, CLIPVisionModel
    if config.clip_config is not None:
        return CLIPVisionModel(CLIPVisionConfig(**config.clip_config))
    return CLIPVisionModel.from_pretrained(config.clip_model_type)


class FlamingoCache:
    """The reference's ``past_key_values`` convention ``(xattn_past, lm_past)`` (modeling_flamingo.py:282-285,303) as an
    object: indexes / unpacks like that 2-tuple, and also answers the few ``Cache`` methods that ``generate()`` of
    transformers >= 5 calls on whatever the model returned (``get_seq_length``, ``reorder_cache``, ``crop``,
    ``batch_repeat_interleave``).  ``xattn`` is a tuple of per-block ``(k, v)``; ``lm`` is whatever the LM returned
    (a ``DynamicCache`` on transformers >= 4.36, legacy tuples before)."""

    is_compileable = False

    def __init__(self, xattn, lm):
        self.xattn = xattn
        self.lm = lm

    def __iter__(self):
        yield self.xattn
        yield self.lm

    def __len__(self):
        return 2

    def __getitem__(self, i):
        return (self.xattn, self.lm)[i]

    def get_seq_length(self, layer_idx: int = 0) -> int:
        if hasattr(self.lm, "get_seq_length"):
            return int(self.lm.get_seq_length())
        return 0 if not self.lm else int(self.lm[0][0].shape[-2])

    def reorder_cache(self, beam_idx: torch.Tensor) -> "FlamingoCache":
        self.xattn = tuple(tuple(t.index_select(0, beam_idx.to(t.device)) for t in kv) for kv in self.xattn)
        if hasattr(self.lm, "reorder_cache"):
            self.lm.reorder_cache(beam_idx)
        else:
            self.lm = tuple(tuple(t.index_select(0, beam_idx.to(t.device)) for t in kv) for kv in self.lm)
        return self

    def crop(self, max_length: int) -> "FlamingoCache":
        """Keep the first `max_length` text positions of the LM cache (the xattn keys/values are per image, not per token)."""
        if hasattr(self.lm, "crop"):
            self.lm.crop(max_length)
        elif self.lm:
            self.lm = tuple((k[:, :, :max_length], v[:, :, :max_length]) for k, v in self.lm)
        return self

    def batch_repeat_interleave(self, repeats: int) -> "FlamingoCache":
        """Every batch row repeated `repeats` times consecutively (beam / candidate expansion), in both caches."""
        self.xattn = tuple(tuple(t.repeat_interleave(repeats, dim=0) for t in kv) for kv in self.xattn)
        if hasattr(self.lm, "batch_repeat_interleave"):
            self.lm.batch_repeat_interleave(repeats)
        elif self.lm:
            self.lm = tuple(tuple(t.repeat_interleave(repeats, dim=0) for t in kv) for kv in self.lm)
        return self


def _repeat_rows(t: torch.Tensor, times: int) -> torch.Tensor:
    """(n, ...) -> (n*times, ...), each row repeated consecutively (beam expansion)."""
    return t.repeat_interleave(times, dim=0)


class FlamingoBaseModel(PreTrainedModel):
    """Shared machinery: builds the vision tower + resampler, splices gated xattn blocks into an LM, runs forward."""

    config_class = FlamingoConfig
    config: FlamingoConfig

    def __init__(self, config: FlamingoConfig, suppress_warnings: bool = True):
        assert isinstance(config, FlamingoConfig)
        super().__init__(config)
        with suppress_model_loading_warnings(suppress_warnings):
            self.vision_encoder = _build_clip(config)
        self.resampler = PerceiverResampler(
            dim=config.dim_visual, depth=config.resampler_depth, dim_head=config.resampler_dim_head,
            heads=config.resampler_heads, num_latents=config.resampler_num_latents,
            num_time_embeds=config.resampler_num_time_embeds, ff_mult=config.resampler_ff_mult,
            act=config.resampler_act)
        # optional callable ids -> embeddings installed by parallel.SplitEmbeddingGrad (data-parallel training only)
        self.embed_lookup = None

    # -- construction helpers -------------------------------------------------------------------------------------
    def _init_layers(self, lm_layers: nn.ModuleList) -> None:
        """Replace every xattn_every-th LM layer by ModifiedLMBlock(layer, ...) (modeling_flamingo.py:76-94)."""
        c = self.config
        for idx in range(0, len(lm_layers), c.xattn_every):
            lm_layers[idx] = ModifiedLMBlock(
                lm_layers[idx], dim=c.dim, dim_visual=c.dim_visual, dim_head=c.xattn_dim_head, heads=c.xattn_heads,
                ff_mult=c.xattn_ff_mult, act=c.xattn_act, n_visual=c.resampler_num_latents)

    def _lm_layers(self) -> nn.ModuleList:
        raise NotImplementedError

    def get_modified_layers(self) -> Iterable[ModifiedLMBlock]:
        return [layer for layer in self._lm_layers() if isinstance(layer, ModifiedLMBlock)]

    # -- freezing / trainable views -------------------------------------------------------------------------------
    def freeze_vm(self) -> None:
        for p in self.vision_encoder.parameters():
            p.requires_grad = False

    def freeze_lm(self) -> None:
        """Freeze the LM except its input embedding (tied with lm_head) and the gated xattn blocks
        (modeling_flamingo.py:105-119)."""
        for p in self.lm.parameters():
            p.requires_grad = False
        self.lm.get_input_embeddings().weight.requires_grad = True
        for layer in self.get_modified_layers():
            for p in layer.xattn_block.parameters():
                p.requires_grad = True

    def unfreeze_lm(self) -> None:
        for p in self.lm.parameters():
            p.requires_grad = True

    def parameters_trainable(self):
        return (p for p in self.parameters() if p.requires_grad)

    def state_dict_trainable(self) -> Dict[str, torch.Tensor]:
        keep = {n for n, p in self.named_parameters() if p.requires_grad}
        return {k: v for k, v in self.state_dict().items() if k in keep}

    # -- vision path ----------------------------------------------------------------------------------------------
    def encode_resample_visuals(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """pixel_values (N c h w) | (b N c h w) | (b N T c h w) -> resampled latents (b N q d)
        (modeling_flamingo.py:140-181)."""
        if pixel_values.ndim == 4:
            b, N, T = 1, pixel_values.shape[0], 1
        elif pixel_values.ndim == 5:
            b, N, T = pixel_values.shape[0], pixel_values.shape[1], 1
        elif pixel_values.ndim == 6:
            b, N, T = pixel_values.shape[:3]
        else:
            raise ValueError("pixel_values must have ndim 5 or 6!")
        flat = pixel_values.reshape(b * N * T, *pixel_values.shape[-3:])
        with torch.no_grad():
            feats = self.vision_encoder(flat).last_hidden_state          # incl. CLS, before post-layernorm
        feats = feats.reshape(b * N, T, feats.shape[-2], feats.shape[-1])
        latents = self.resampler(feats)                                   # (b*N, q, d): frames are folded into keys
        return latents.reshape(b, N, latents.shape[-2], latents.shape[-1])


    # -- loss head ----------------------------------------------------------------------------------------------------
    def _lm_head_logits(self, hidden: torch.Tensor):
        """logits = lm_head(hidden) (modeling_flamingo.py:279).  The <EOC> resize makes the vocabulary 50 258 / 50 273
        wide, which pushes cuBLAS onto its unaligned (4x slower) bf16 kernels; on CUDA half-precision inputs the tied
        weight is therefore zero-padded to a multiple of 64 rows and the padded columns get a -inf bias, so softmax /
        cross-entropy over the padded logits are bit-for-bit the loss over the real vocabulary.  Returns
        (logits[..., :vocab] view, padded logits or None)."""
        w = self.lm_head.weight
        vocab = w.shape[0]
        pad = (-vocab) % 64
        if pad == 0 or not hidden.is_cuda or hidden.dtype not in (torch.bfloat16, torch.float16) or self.lm_head.bias is not None:
            return self.lm_head(hidden), None
        if torch.is_grad_enabled() and w.requires_grad:
            # training: the padded weight must stay a differentiable function of the (tied, trainable) embedding
            bias = hidden.new_zeros(vocab + pad)
            bias[vocab:] = float("-inf")
            padded = F.linear(hidden, F.pad(w, (0, 0, 0, pad)), bias)
            return padded[..., :vocab], padded
        # inference / generation: pad once per weight version instead of copying vocab x D on every decoded token
        key = (w._version, w.data_ptr(), w.device, w.dtype, hidden.dtype)
        cached = getattr(self, "_padded_head", None)
        if cached is None or cached[0] != key:
            bias = torch.zeros(vocab + pad, dtype=hidden.dtype, device=w.device)
            bias[vocab:] = float("-inf")
            cached = (key, F.pad(w.detach(), (0, 0, 0, pad)).to(hidden.dtype), bias)
            self._padded_head = cached
        padded = F.linear(hidden, cached[1], cached[2])
        return padded[..., :vocab], padded

    def _loss(self, logits, padded, labels, reduction):
        """Shifted next-token loss; with FlamingoConfig.fused_cross_entropy through the library's row kernels."""
        if (getattr(self.config, "fused_cross_entropy", False) and reduction == "mean" and logits.is_cuda
                and logits.dtype == torch.bfloat16):
            from . import functional as Fn
            full = padded if padded is not None else logits
            if full.size(-1) % 8 == 0 and full.is_contiguous():
                # the last position of every sequence is ignored instead of copying the shifted logits
                tgt = torch.full_like(labels, -100)
                tgt[..., :-1] = labels[..., 1:]
                return Fn.cross_entropy(full.reshape(-1, full.size(-1)), tgt.reshape(-1), logits.size(-1), -100)
        return self._shifted_cross_entropy(logits, padded, labels, reduction)

    @staticmethod
    def _shifted_cross_entropy(logits, padded, labels, reduction):
        vocab = logits.size(-1)
        if padded is None or reduction != "mean":
            return F.cross_entropy(logits[..., :-1, :].reshape(-1, vocab), labels[..., 1:].reshape(-1), reduction=reduction)
        # same mean over the B*(S-1) shifted positions, without copying the logits: the last position is ignored
        tgt = torch.full_like(labels, -100)
        tgt[..., :-1] = labels[..., 1:]
        return F.cross_entropy(padded.reshape(-1, padded.size(-1)), tgt.reshape(-1), ignore_index=-100, reduction="mean")

    # -- forward ----------------------------------------------------------------------------------------------------
    def forward(self, input_ids=None, attention_mask=None, media_locations=None, pixel_values=None,
                visual_features=None, head_mask=None, inputs_embeds=None, use_cache: bool = False,
                past_key_values=None, return_dict: bool = True, labels=None, loss_reduction: str = "mean", **kwargs
                ) -> CausalLMOutputWithPast:
        assert return_dict, "can only use return_dict=True at the moment!"
        assert (input_ids is None) != (inputs_embeds is None), "you must pass either input_ids or inputs_embeds!"
        ref = input_ids if input_ids is not None else inputs_embeds
        batch_size, seq_length = ref.shape[:2]
        device = ref.device
        xattn_past = None if past_key_values is None else past_key_values[0]
        lm_past = None if past_key_values is None else past_key_values[1]

        if visual_features is None:
            if xattn_past is None and pixel_values is not None:
                assert pixel_values.size(0) == batch_size, "pixel_values must have the same batch size as the textual input!"
                visual_features = self.encode_resample_visuals(pixel_values)
            else:   # shape-only dummy: keys/values come from the cache, or there is no image at all
                visual_features = torch.zeros((batch_size, 1, self.config.resampler_num_latents, self.config.dim_visual),
                                              dtype=torch.float32, device=device)
        if media_locations is None:
            media_locations = torch.zeros((batch_size, seq_length), dtype=torch.int, device=device)

        modified = list(self.get_modified_layers())
        # text_time = cumsum(media_locations) (gated_cross_attention.py:97) depends only on media_locations: once per forward
        text_time = None
        if modified and media_locations.is_cuda and hasattr(modified[0].xattn_block, "_fp"):
            from . import functional as Fn
            text_time = Fn.text_time_of(media_locations)
        for i, layer in enumerate(modified):
            if text_time is None:
                layer.condition(visual_features, media_locations, None if xattn_past is None else xattn_past[i])
            else:
                layer.condition(visual_features, media_locations, None if xattn_past is None else xattn_past[i], text_time=text_time)

        if input_ids is not None and self.embed_lookup is not None and torch.is_grad_enabled():
            # data-parallel training: the lookup's weight gradient is exchanged sparsely (parallel.SplitEmbeddingGrad)
            inputs_embeds, input_ids = self.embed_lookup(input_ids), None
        lm_kwargs = dict(input_ids=input_ids, attention_mask=attention_mask, inputs_embeds=inputs_embeds,
                         use_cache=use_cache, past_key_values=lm_past, return_dict=True, **kwargs)
        if head_mask is not None:
            lm_kwargs["head_mask"] = head_mask
        out = self.lm(**lm_kwargs)
        logits, padded = self._lm_head_logits(out.last_hidden_state)

        xattn_kv = tuple(layer.kv_output for layer in modified) if use_cache else None

        loss = None
        if labels is not None:   # next-token loss: positions < n predict n (modeling_flamingo.py:287-298)
            loss = self._loss(logits, padded, labels, loss_reduction)
        return CausalLMOutputWithPast(
            loss=loss, logits=logits,
            past_key_values=FlamingoCache(xattn_kv, out.past_key_values) if use_cache else None,
            hidden_states=getattr(out, "hidden_states", None), attentions=getattr(out, "attentions", None))


def _fuse_gelu_new(lm: nn.Module) -> int:
    """Replace HF's ``NewGELUActivation`` modules (tanh-GELU spelled out as ~8 elementwise torch ops, each a kernel launch
    and an HBM round trip of the 4*D-wide hidden state, forward and backward) by ``nn.GELU(approximate="tanh")`` — the
    same formula in one kernel.  Host-side PyTorch only; parameters and checkpoint keys are untouched."""
    from transformers.activations import NewGELUActivation
    n = 0
    for module in lm.modules():
        for name, child in list(module.named_children()):
            if isinstance(child, NewGELUActivation):
                setattr(module, name, nn.GELU(approximate="tanh"))
                n += 1
    return n


class FlamingoGPT2(FlamingoBaseModel):
    def __init__(self, config: FlamingoConfig):
        from transformers import GPT2Config, GPT2LMHeadModel
        assert config.lm.startswith("gpt")
        super().__init__(config)
        if config.lm_config is not None:
            base_lm = GPT2LMHeadModel(GPT2Config(**config.lm_config))
        else:
            base_lm = GPT2LMHeadModel.from_pretrained(config.lm)
        assert config.dim == base_lm.config.n_embd, \
            f"specified {config.dim=} in FlamingoConfig, but {config.lm} has hidden size={base_lm.config.n_embd}"
        base_lm.resize_token_embeddings(base_lm.config.vocab_size + 1)      # <EOC>
        if getattr(config, "lm_fused_gelu", True):
            _fuse_gelu_new(base_lm)
        self.lm = base_lm.transformer
        self.lm_head = base_lm.lm_head
        self._init_layers(self.lm.h)

    def _lm_layers(self):
        return self.lm.h


class FlamingoOPT(FlamingoBaseModel):
    def __init__(self, config: FlamingoConfig):
        from transformers import OPTConfig, OPTForCausalLM
        assert config.lm.startswith("facebook/opt")
        super().__init__(config)
        if config.lm_config is not None:
            base_lm = OPTForCausalLM(OPTConfig(**config.lm_config))
        else:
            base_lm = OPTForCausalLM.from_pretrained(config.lm)
        assert config.dim == base_lm.config.hidden_size, \
            f"specified {config.dim=} in FlamingoConfig, but {config.lm} has hidden size={base_lm.config.hidden_size}"
        base_lm.resize_token_embeddings(base_lm.config.vocab_size + 1)
        self.lm = base_lm.model
        self.lm_head = base_lm.lm_head
        self._init_layers(self.lm.decoder.layers)

    def _lm_layers(self):
        return self.lm.decoder.layers


class FlamingoModel(PreTrainedModel, GenerationMixin):
    """LM-agnostic front: picks FlamingoGPT2 / FlamingoOPT from ``config.lm`` and forwards to it
    (modeling_flamingo.py:359-712)."""

    config_class = FlamingoConfig
    config: FlamingoConfig
    _LANGUAGE_MODEL_VERSIONS = {"gpt2": FlamingoGPT2, "facebook/opt": FlamingoOPT}
    _keys_to_ignore_on_load_missing = [r"flamingo.vision_encoder"]

    def __init__(self, config: FlamingoConfig, model_class: Optional[type] = None):
        super().__init__(config)
        if model_class is None:
            model_class = self._find_flamingo_class(config.lm)
        self.flamingo: FlamingoBaseModel = model_class(config)
        # beam search of transformers >= 5 reads config.get_text_config().vocab_size
        config.vocab_size = self.flamingo.lm_head.weight.shape[0]
        if config.freeze_language_model:
            self.freeze_lm()
        if config.freeze_vision_model:
            self.freeze_vm()

    @classmethod
    def is_lm_supported(cls, lm_id: str) -> bool:
        return any(lm_id.startswith(prefix) for prefix in cls._LANGUAGE_MODEL_VERSIONS)

    @classmethod
    def _find_flamingo_class(cls, language_model_id: str):
        for prefix, klass in cls._LANGUAGE_MODEL_VERSIONS.items():
            if language_model_id.startswith(prefix):
                return klass
        raise ValueError(f"unsupported language model {language_model_id}")

    def parameters_trainable(self):
        return self.flamingo.parameters_trainable()

    def freeze_vm(self):
        self.flamingo.freeze_vm()

    def freeze_lm(self):
        self.flamingo.freeze_lm()

    def unfreeze_lm(self):
        self.flamingo.unfreeze_lm()

    def state_dict_trainable(self):
        return self.flamingo.state_dict_trainable()

    def forward(self, input_ids=None, attention_mask=None, media_locations=None, pixel_values=None,
                visual_features=None, head_mask=None, inputs_embeds=None, use_cache: bool = False,
                past_key_values=None, return_dict: bool = True, labels=None, loss_reduction: str = "mean", **kwargs
                ) -> CausalLMOutputWithPast:
        return self.flamingo(input_ids=input_ids, attention_mask=attention_mask, media_locations=media_locations,
                             pixel_values=pixel_values, visual_features=visual_features, head_mask=head_mask,
                             inputs_embeds=inputs_embeds, use_cache=use_cache, past_key_values=past_key_values,
                             return_dict=return_dict, labels=labels, loss_reduction=loss_reduction, **kwargs)

    # -- generation plumbing ------------------------------------------------------------------------------------------
    @classmethod
    def _supports_default_dynamic_cache(cls) -> bool:
        """generate() of transformers >= 5 must not pre-build a DynamicCache from this (composite) config: the first
        forward returns a FlamingoCache holding both the xattn K/V and the LM's own cache."""
        return False

    def prepare_inputs_for_generation(self, input_ids, media_locations=None, attention_mask=None, pixel_values=None,
                                      visual_features=None, past=None, past_key_values=None, next_sequence_length=None,
                                      is_first_iteration=None, position_ids=None, **kwargs) -> Dict[str, Any]:
        """Expand visual inputs / media_locations to the (beam-expanded) text batch and keep only the last token
        once a cache exists (modeling_flamingo.py:464-523).  ``next_sequence_length`` / ``is_first_iteration`` /
        ``position_ids`` are bookkeeping that generate() of transformers >= 5 passes; position ids are cut to the
        tokens actually fed."""
        n = input_ids.shape[0]

        def fit(t):
            if t is None or t.shape[0] == n:
                return t
            assert n % t.shape[0] == 0
            return _repeat_rows(t, n // t.shape[0])

        cache = past_key_values if past_key_values is not None else past
        if cache is not None:
            input_ids = input_ids[:, -1:]
        if position_ids is not None:
            kwargs["position_ids"] = position_ids[..., -input_ids.shape[1]:]
        return dict(input_ids=input_ids, past_key_values=cache, media_locations=fit(media_locations),
                    attention_mask=attention_mask, pixel_values=fit(pixel_values), visual_features=fit(visual_features),
                    **kwargs)

    def _reorder_cache(self, past, beam_idx):
        """Reorder both caches for beam search (modeling_flamingo.py:525-548)."""
        if isinstance(past, FlamingoCache):
            return past.reorder_cache(beam_idx)
        xattn_past, lm_past = past

        def pick(layer):
            return tuple(t.index_select(0, beam_idx.to(t.device)) for t in layer)

        xattn_beam = tuple(pick(layer) for layer in xattn_past)
        if hasattr(lm_past, "reorder_cache"):          # transformers >= 4.36 Cache objects
            lm_past.reorder_cache(beam_idx)
            return xattn_beam, lm_past
        return xattn_beam, tuple(pick(layer) for layer in lm_past)

    @torch.no_grad()
    def generate_captions(self, processor, pixel_values=None, images=None, prompt: str = "<image>", max_length: int = 150,
                          num_beams: int = 1, device=None, **kwargs):
        """Caption a batch of images with the same prompt (modeling_flamingo.py:550-605)."""
        device = self.device if device is None else device
        if images is not None:
            assert pixel_values is None, "you can only pass either images or visual features to generate_captions()!"
            if not isinstance(images, (list, tuple)):
                images = [images]
            pixel_values = processor(images=images, device=device)["pixel_values"]
        assert pixel_values is not None, "you must pass either images or visual features to generate_captions()!"
        n = pixel_values.size(0)
        ids, media_locations, mask = processor.encode_text(prompt, device)
        lm_cfg = self.flamingo.lm.config
        out_ids = self.generate(
            inputs=ids[:1].expand(n, -1), media_locations=media_locations[:1].expand(n, -1),
            attention_mask=mask[:1].expand(n, -1), pixel_values=pixel_values, num_beams=num_beams, early_stopping=True,
            use_cache=True, bos_token_id=lm_cfg.bos_token_id, eos_token_id=lm_cfg.eos_token_id,
            pad_token_id=lm_cfg.eos_token_id, max_length=max_length, **kwargs)
        texts = processor.tokenizer.batch_decode(out_ids, skip_special_tokens=True)
        return [processor.remove_tags(t) for t in texts]

    @torch.no_grad()
    def score_sequences(self, input_ids, media_locations, attention_mask, pixel_values=None, visual_features=None,
                        k: int = 100000) -> torch.Tensor:
        """EXPERIMENTAL zero-shot scoring (modeling_flamingo.py:607-712): log-prob of each candidate sequence given
        the same visual input; the common prefix is run once and its caches are shared by the top-k continuations."""
        assert visual_features is None or visual_features.ndim == 3, "visual_features must have shape (N q d)"
        n_choices = input_ids.size(0)
        n_reuse = get_common_prefix_length(input_ids)
        k = min(k, n_choices)
        first = self.flamingo(
            input_ids=input_ids[:1, :n_reuse], media_locations=media_locations[:1, :n_reuse],
            attention_mask=attention_mask[:1, :n_reuse],
            pixel_values=None if pixel_values is None else pixel_values.unsqueeze(0),
            visual_features=None if visual_features is None else visual_features.unsqueeze(0), use_cache=True)
        next_tokens = input_ids[:, n_reuse]
        top = first.logits[0, -1, :].index_select(0, next_tokens).topk(k).indices
        xattn_past = [tuple(t.expand(k, *t.shape[1:]) for t in kv) for kv in first.past_key_values[0]]
        lm_cache = first.past_key_values[1]
        if hasattr(lm_cache, "batch_repeat_interleave"):     # Cache object: drop the last prefix position, expand
            lm_cache.crop(n_reuse - 1)
            lm_cache.batch_repeat_interleave(k)
            lm_past = lm_cache
        else:
            lm_past = [(ks.expand(k, *ks.shape[1:])[:, :, :-1, :], vs.expand(k, *vs.shape[1:])[:, :, :-1, :])
                       for ks, vs in lm_cache]
        choice_ids = input_ids[top, n_reuse - 1:]
        second = self.flamingo(input_ids=choice_ids, media_locations=media_locations[top],
                               attention_mask=attention_mask[top], pixel_values=None, visual_features=None,
                               past_key_values=(xattn_past, lm_past), labels=choice_ids, loss_reduction="none")
        losses = second.loss.reshape(k, -1).sum(dim=1)
        scores = torch.full([n_choices], torch.finfo(torch.float).min, device=losses.device)
        scores[top] = -losses
        return scores.detach()
