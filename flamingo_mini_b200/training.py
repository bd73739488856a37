"""Minimal training driver for the hot path: what the reference gets from HuggingFace ``Trainer`` in
training/train.py:95-190 + train.sh (AdamW with the Trainer defaults, DDP under torchrun, COCO caption collator,
trainable-only checkpoints), restated over this package's flat parameter buffers:

* ``ArenaAdamW`` — every hot-path module keeps its parameters in ONE flat fp32 buffer and its backward writes ONE flat
  gradient arena (functional.FlatParams), so the optimizer step is a single fused AdamW update per module instead of one
  per tensor; the bf16 tensor-core shadow is invalidated afterwards.  Parameters outside the modules (the tied token
  embedding the reference keeps trainable, modeling_flamingo.py:115) take the ordinary per-tensor path.
* ``DataCollator`` — (image tensor, caption) pairs -> model kwargs, as training/train.py:71-84.
* ``save_trainable`` / ``load_trainable`` — ``state_dict_trainable`` checkpoints (modeling_flamingo.py:125-130).
* ``train`` — forward / backward / gradient exchange (parallel.GradArenaReducer) / optimizer step.

Host-side PyTorch only; the arithmetic of the two modules stays in the sm_100a library.
"""
from __future__ import annotations

import math
import os
from typing import Callable, Dict, Iterable, List, Optional

import torch
from torch import nn

from . import functional as Fn
from .parallel import GradArenaReducer, hot_path_modules


class ArenaAdamW:
    """AdamW (decoupled weight decay; defaults = transformers.TrainingArguments: lr 5e-5, betas (0.9, 0.999), eps 1e-8,
    weight_decay 0) over flat buffers.  On CUDA arenas the whole step of a module is ONE launch of the library's fused kernel
    (fm_adamw_step: decay, both moments, bias-corrected update, gradient-clip scale and the bf16 tensor-core shadow of the new
    parameters in a single pass over HBM).  Weight decay, when non-zero, is grouped exactly like HF
    ``Trainer.get_decay_parameter_names``: every parameter EXCEPT those of ``nn.LayerNorm`` modules and names containing
    "bias" is decayed — i.e. the gates, ``latents`` and ``time_pos_emb`` ARE decayed, as they are under the reference's
    recipe (training/train.py uses the stock Trainer)."""

    def __init__(self, modules: Iterable[nn.Module], extra_params: Iterable[nn.Parameter] = (), lr: float = 5e-5,
                 betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        self.modules = list(modules)
        self.extra = [p for p in extra_params if p.requires_grad]
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.step_count = 0
        self._state: Dict[int, Dict[str, torch.Tensor]] = {}
        self._decay_mask: Dict[int, torch.Tensor] = {}
        self._extra_opt = (torch.optim.AdamW(self.extra, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
                           if self.extra else None)

    # -- per-module flat state --------------------------------------------------------------------------------------
    def _flat_state(self, mod):
        flat = mod._fp.ensure()
        st = self._state.get(id(mod))
        if st is None or st["m"].device != flat.device:
            st = {"m": torch.zeros_like(flat), "v": torch.zeros_like(flat)}
            self._state[id(mod)] = st
        return flat, st

    def _mask(self, mod, flat):
        """1.0 where weight decay applies, 0.0 for the parameters of nn.LayerNorm modules and names containing "bias"
        (HF Trainer.get_decay_parameter_names = get_parameter_names(model, [nn.LayerNorm]) minus "bias" names)."""
        mk = self._decay_mask.get(id(mod))
        if mk is None or mk.device != flat.device:
            mk = torch.zeros_like(flat)
            norm_params = {id(p) for m in mod.modules() if isinstance(m, nn.LayerNorm) for p in m.parameters()}
            names = {id(p): n for n, p in mod.named_parameters()}
            for p, off in mod._fp.slots:
                if id(p) not in norm_params and "bias" not in names.get(id(p), ""):
                    mk[off:off + p.numel()] = 1.0
            self._decay_mask[id(mod)] = mk
        return mk

    @torch.no_grad()
    def step(self, lr: Optional[float] = None, grad_scale: Optional[torch.Tensor] = None) -> None:
        """grad_scale: optional 0-dim device tensor multiplied into every gradient (clipping) inside the update itself."""
        from . import _lib
        lr = self.lr if lr is None else lr
        self.step_count += 1
        b1, b2 = self.betas
        bc1 = 1.0 - b1 ** self.step_count
        bc2 = 1.0 - b2 ** self.step_count
        for mod in self.modules:
            if getattr(mod, "_last_grad_arena", None) is None:
                continue                       # module did not take part in this step's backward
            g, _ = mod._fp.current_grad(mod)   # the cached arena the .grads alias (after accumulation: the accumulated sum)
            if g is None:
                continue
            flat, st = self._flat_state(mod)
            if flat.is_cuda and _lib.has("fm_adamw_step"):
                fp = mod._fp
                if fp._shadow is None or fp._shadow.device != flat.device:
                    fp._shadow = torch.empty(fp.total, dtype=torch.bfloat16, device=flat.device)
                mask = self._mask(mod, flat) if self.weight_decay != 0.0 else None
                gs = None if grad_scale is None else grad_scale.to(torch.float32).reshape(1).contiguous()
                _lib.check(_lib.load().fm_adamw_step(Fn._ptr(flat), Fn._ptr(g), Fn._ptr(st["m"]), Fn._ptr(st["v"]), Fn._ptr(fp._shadow),
                                                     Fn._ptr(mask), Fn._ptr(gs), fp.total, lr, b1, b2, self.eps, self.weight_decay,
                                                     self.step_count, Fn._stream()), "fm_adamw_step")
                fp._shadow_ver = sum(p._version for p, _ in fp.slots)       # the kernel wrote the up-to-date bf16 shadow
                continue
            if grad_scale is not None:
                g = g * grad_scale.to(g.dtype)
            if self.weight_decay != 0.0:
                flat.addcmul_(flat, self._mask(mod, flat), value=-lr * self.weight_decay)
            st["m"].lerp_(g, 1.0 - b1)
            st["v"].mul_(b2).addcmul_(g, g, value=1.0 - b2)
            denom = (st["v"].sqrt() / math.sqrt(bc2)).add_(self.eps)
            flat.addcdiv_(st["m"], denom, value=-lr / bc1)
            mod._fp.invalidate_shadow()
        if self._extra_opt is not None:
            if grad_scale is not None:
                for p in self.extra:
                    if p.grad is not None:
                        p.grad.mul_(grad_scale.to(p.grad.dtype))
            for grp in self._extra_opt.param_groups:
                grp["lr"] = lr
            self._extra_opt.step()

    def zero_grad(self) -> None:
        for mod in self.modules:
            mod._last_grad_arena = None
            for p in mod.parameters():
                p.grad = None
        for p in self.extra:
            p.grad = None

    def state_dict(self) -> dict:
        return {"step": self.step_count,
                "modules": [self._state.get(id(m)) for m in self.modules],
                "extra": self._extra_opt.state_dict() if self._extra_opt is not None else None}

    def load_state_dict(self, sd: dict) -> None:
        self.step_count = sd["step"]
        for m, st in zip(self.modules, sd["modules"]):
            if st is not None:
                flat = m._fp.ensure()
                self._state[id(m)] = {k: v.to(flat.device) for k, v in st.items()}
        if self._extra_opt is not None and sd.get("extra") is not None:
            self._extra_opt.load_state_dict(sd["extra"])


def constant_schedule_with_warmup(base_lr: float, warmup_steps: int) -> Callable[[int], float]:
    """lr used by the optimizer step number `step` (0-based) under transformers.get_constant_schedule_with_warmup (the scheduler
    named in training/train.py:165): lr = base * step / max(1, warmup) while step < warmup — the very first update runs at lr 0,
    exactly as under HF — and base afterwards."""
    return lambda step: base_lr * (float(step) / float(max(1, warmup_steps)) if step < warmup_steps else 1.0)


class DataCollator:
    """(pixel_values, caption) pairs -> model kwargs (training/train.py:71-84): labels = input_ids."""

    def __init__(self, processor):
        self.processor = processor

    def __call__(self, batch):
        pixel_values, sentences = zip(*batch)
        inputs = self.processor(text=list(sentences))
        return dict(pixel_values=torch.stack(list(pixel_values)), labels=inputs["input_ids"], **inputs)


def save_trainable(model: nn.Module, path: str, optimizer: Optional[ArenaAdamW] = None, step: int = 0) -> None:
    """Checkpoint of the trainable parameters only (modeling_flamingo.py:125-130) + optimizer state."""
    base = getattr(model, "flamingo", model)
    sd = {k: v.detach().cpu().clone() for k, v in base.state_dict_trainable().items()}
    tmp = f"{path}.tmp.{os.getpid()}"
    torch.save({"trainable": sd, "step": step, "optimizer": None if optimizer is None else optimizer.state_dict()}, tmp)
    os.replace(tmp, path)


def load_trainable(model: nn.Module, path: str, optimizer: Optional[ArenaAdamW] = None) -> int:
    base = getattr(model, "flamingo", model)
    ck = torch.load(path, map_location="cpu")
    res = base.load_state_dict(ck["trainable"], strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    for m in hot_path_modules(model):
        m._fp.invalidate_shadow()
    if optimizer is not None and ck.get("optimizer") is not None:
        optimizer.load_state_dict(ck["optimizer"])
    return int(ck.get("step", 0))


def train(model: nn.Module, batches: Iterable[dict], steps: int, lr: float = 5e-5, warmup_steps: int = 0,
          weight_decay: float = 0.0, max_grad_norm: Optional[float] = None, log_every: int = 0,
          split_embedding: bool = False) -> List[float]:
    """Runs `steps` optimisation steps; returns the per-step losses.  Under torchrun (torch.distributed initialised) the
    trainable gradients are averaged over ranks exactly once per step (reference: implicit DDP, train.sh:26,36)."""
    import torch.distributed as dist
    hot = hot_path_modules(model)
    hot_ids = {id(p) for m in hot for p in m.parameters()}
    extra = [p for p in model.parameters() if p.requires_grad and id(p) not in hot_ids]
    world = dist.get_world_size() if dist.is_initialized() else 1
    reducer = GradArenaReducer(hot, extra_params=extra) if world > 1 else None
    if reducer is not None and split_embedding:
        from .parallel import SplitEmbeddingGrad
        SplitEmbeddingGrad.install(model, reducer)
    opt = ArenaAdamW(hot, extra, lr=lr, weight_decay=weight_decay)
    sched = constant_schedule_with_warmup(lr, warmup_steps)
    losses: List[float] = []
    model.train()
    it = iter(batches)
    for step in range(steps):
        batch = next(it)
        opt.zero_grad()
        out = model(**batch)
        out.loss.backward()
        if Fn._PENDING:                  # only with the defer_join switch on
            Fn.side_join()
        if reducer is not None:
            reducer.finish()
        scale = None
        if max_grad_norm is not None:      # torch.nn.utils.clip_grad_norm_ semantics; the scale is applied inside the optimizer step
            grads = [g for g in (m._fp.current_grad(m)[0] for m in hot) if g is not None] + [p.grad for p in extra if p.grad is not None]
            total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g.float()) for g in grads]))
            scale = (max_grad_norm / (total + 1e-6)).clamp(max=1.0)
        opt.step(sched(step), grad_scale=scale)
        losses.append(float(out.loss.detach()))
        if log_every and (step + 1) % log_every == 0 and (not dist.is_initialized() or dist.get_rank() == 0):
            print(f"step {step + 1}: loss {losses[-1]:.4f} lr {sched(step):.2e}", flush=True)
    if reducer is not None:
        reducer.detach()
    return losses
