"""Data parallelism for the hot path: one process per GPU, parameters replicated, the batch sharded by rank, and ONE
exchange step — an all-reduce (mean) of the trainable gradients — which is what the reference gets implicitly from
DistributedDataParallel under torchrun (training/train.sh:26,36 of the reference).

Every hot-path module writes its parameter gradients into one flat fp32 arena (functional.FlatParams), so the
exchange is one collective per module instead of one per tensor.  The arena of block i is complete the moment that
block's backward kernels are enqueued (backward visits the LM top-down, so blocks finish in the order L-1 .. 0 and
the resampler last); ``GradArenaReducer`` launches the all-reduce right there, asynchronously, so it overlaps the
backward of the remaining blocks and of the frozen LM.  On CUDA the collective is NCCL over NVLink/NVSwitch; on CPU
tensors (tests) the same code runs over gloo.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def hot_path_modules(model: torch.nn.Module) -> List[torch.nn.Module]:
    """Every module that owns a flat gradient arena (PerceiverResampler, GatedCrossAttentionBlock)."""
    return [m for m in model.modules() if hasattr(m, "_fp") and hasattr(m, "_grad_ready_hook")]


class GradArenaReducer:
    def __init__(self, modules: Iterable[torch.nn.Module], extra_params: Iterable[torch.nn.Parameter] = (),
                 group: Optional[dist.ProcessGroup] = None):
        self.modules = list(modules)
        self.extra_params = [p for p in extra_params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._pending = []
        self.bytes_reduced = 0
        for m in self.modules:
            m._grad_ready_hook = self._on_arena_ready

    def detach(self) -> None:
        for m in self.modules:
            m._grad_ready_hook = None

    # called from inside the module's backward (autograd thread), right after its kernels were enqueued
    def _on_arena_ready(self, module, arena: torch.Tensor) -> None:
        if self.world == 1:
            return
        self._launch(arena)

    def _launch(self, t: torch.Tensor) -> None:
        self.bytes_reduced += t.numel() * t.element_size()
        if t.is_cuda:
            work = dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
            self._pending.append((work, None))
        else:   # gloo has no AVG
            work = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._pending.append((work, t))

    def finish(self) -> None:
        """Call after loss.backward(): reduces the remaining (non-arena) trainable gradients, e.g. the token embedding
        the reference keeps trainable (modeling_flamingo.py:115), and waits for every outstanding collective."""
        if self.world > 1:
            for p in self.extra_params:
                if p.grad is not None:
                    self._launch(p.grad)
        for work, scale_me in self._pending:
            work.wait()
            if scale_me is not None:
                scale_me.div_(self.world)
        self._pending.clear()


def shard_batch(global_batch: int, rank: int, world: int) -> range:
    """Contiguous shard of sample indices owned by ``rank`` (weak scaling keeps per-rank size fixed instead)."""
    per = (global_batch + world - 1) // world
    return range(min(rank * per, global_batch), min((rank + 1) * per, global_batch))
