"""FlamingoProcessor — tokenizer + CLIP image processor wrapper with the reference's interface
(flamingo_mini/flamingo_processor.py:11-147): ``encode_text``, ``prepare_caption(s)``, ``remove_tags``,
``get_media_locations``, ``preprocess_images`` and ``__call__(images, text, device)``.

CPU-side preprocessing, outside the CUDA hot path.  The HuggingFace tokenizer / image processor are fetched with
``from_pretrained`` unless ready-made objects are passed in (the build machines have no hub access).
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from .configuration_flamingo import FlamingoConfig


class FlamingoProcessor:
    def __init__(self, config: FlamingoConfig, use_fast: bool = True, eoc_token: str = "<EOC>", tokenizer=None,
                 vision_processor=None):
        self.config = config
        self.eoc_token = eoc_token
        if vision_processor is None:
            from transformers import CLIPImageProcessor
            vision_processor = CLIPImageProcessor.from_pretrained(config.clip_model_type)
        self.vision_processor = vision_processor
        if tokenizer is None:
            from transformers import AutoTokenizer
            if config.lm.startswith("gpt2"):
                tokenizer = AutoTokenizer.from_pretrained("gpt2", use_fast=use_fast)
            elif config.lm.startswith("facebook/opt"):
                tokenizer = AutoTokenizer.from_pretrained("facebook/opt-30b", use_fast=use_fast)
            else:
                raise ValueError(f"unsupported language model {config.lm}")
        self.tokenizer = tokenizer
        self.tokenizer.add_bos_token = True
        self.tokenizer.pad_token = self.tokenizer.eos_token
        self.tokenizer.add_tokens(self.eoc_token)
        # "<image>" is located through its first character: '<' tokenises differently after a space
        self.leq_ids = [self.tokenizer.encode("<")[-1], self.tokenizer.encode(" <")[-1]]

    # ------------------------------------------------------------------ text
    def encode_text(self, text: str | List[str], device: torch.device | None = None, max_length=None, length=None,
                    return_tensors="pt", return_attention_mask=True) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        kw = dict(return_tensors=return_tensors)
        if length is not None:
            kw.update(return_attention_mask=return_attention_mask, padding="max_length", truncation=True, max_length=length)
        elif max_length is not None:
            kw.update(return_attention_mask=return_attention_mask, padding=True, truncation=True, max_length=max_length)
        else:
            kw.update(padding=True)
        enc = self.tokenizer(text, **kw)
        media_locations = self.get_media_locations(enc.input_ids)
        return enc.input_ids.to(device), media_locations.to(device), enc.attention_mask.to(device)

    def get_media_locations(self, input_ids: torch.Tensor) -> torch.Tensor:
        """1 where a token is the '<' that opens an <image> tag (int64, like the reference)."""
        hits = torch.zeros_like(input_ids, dtype=torch.int64)
        for tok in self.leq_ids:
            hits = hits + (input_ids == tok).to(torch.int64)
        return hits

    def prepare_caption(self, caption: str) -> str:
        return "<image>" + caption + self.eoc_token + self.tokenizer.eos_token

    def prepare_captions(self, captions: List[str]) -> List[str]:
        return [self.prepare_caption(c) for c in captions]

    def _remove_tags(self, text: str) -> str:
        for tag in ("<image>", self.tokenizer.eos_token, self.eoc_token, self.tokenizer.pad_token):
            text = text.replace(tag, "")
        return text.strip()

    def remove_tags(self, text: str | List[str]):
        return self._remove_tags(text) if isinstance(text, str) else [self._remove_tags(t) for t in text]

    # ------------------------------------------------------------------ images
    def preprocess_images(self, images):
        return self.vision_processor(images=images, return_tensors="pt", padding=True)

    def __call__(self, images=None, text: str | List[str] | None = None, device: torch.device | None = None):
        out = {}
        if images is not None:
            out["pixel_values"] = self.preprocess_images(images)["pixel_values"].to(device)
        if text is not None:
            ids, media_locations, mask = self.encode_text(text, device=device)
            out.update(input_ids=ids, media_locations=media_locations, attention_mask=mask)
        return out
