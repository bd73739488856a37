"""Data parallelism for the hot path: one process per GPU, parameters replicated, the batch sharded by rank, and ONE
exchange step — an all-reduce (mean) of the trainable gradients — which is what the reference gets implicitly from
DistributedDataParallel under torchrun (training/train.sh:26,36 of the reference).

Every hot-path module writes its parameter gradients into one flat fp32 arena (functional.FlatParams), so the
exchange is one collective per module instead of one per tensor.  The arena of block i is complete the moment that
block's backward kernels are enqueued (backward visits the LM top-down, so blocks finish in the order L-1 .. 0 and
the resampler last); ``GradArenaReducer`` launches the all-reduce right there, asynchronously, so it overlaps the
backward of the remaining blocks and of the frozen LM.  On CUDA the collective is NCCL over NVLink/NVSwitch; on CPU
tensors (tests) the same code runs over gloo.

Gradient accumulation: call ``finish()`` after every backward.  A backward that adds into existing ``.grad`` is reduced in
``finish()`` over the accumulated gradient (see ``_accumulating``); ``no_sync()`` skips the exchange for the micro-batches
inside it, as DistributedDataParallel.no_sync() does.
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
                 group: Optional[dist.ProcessGroup] = None, per_layer: bool = False, wire_dtype: Optional[torch.dtype] = None,
                 bucket_blocks: int = 1):
        """per_layer=True: modules whose backward can report per-layer completion (PerceiverResampler through
        fm_resampler_bwd_notify) get their arena reduced layer by layer while backward is still running, so only the last
        layer's slice is left in the exposed tail.
        wire_dtype=torch.bfloat16: fp32 CUDA arenas cross NVLink as bf16 (cast, all-reduce(AVG), cast back into the fp32 arena
        in finish()) — the same trade as DDP's bf16_compress_hook: half the bytes on the wire and half the time NCCL's CTAs
        sit on SMs the GEMMs want, for one bf16 rounding (2^-9 relative) of every averaged gradient element."""
        self.modules = list(modules)
        self.per_layer = per_layer
        self.wire_dtype = wire_dtype
        # bucket_blocks = K > 1: the gradient arenas of K consecutive gated xattn blocks are slices of ONE buffer and are reduced
        # by one collective when the K-th of them (in backward order) is complete: fewer, larger all-reduces
        self.bucket_blocks = max(1, int(bucket_blocks))
        self._buckets = {}           # id(module) -> (bucket index, position); see _build_buckets
        self._bucket_state = []
        self.extra_params = [p for p in extra_params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._pending = []
        self._deferred = []          # modules whose backward ACCUMULATED into existing .grad: reduced in finish(), see _accumulating
        self._sync = True            # False inside no_sync()
        self._split: Optional["SplitEmbeddingGrad"] = None
        self.bytes_reduced = 0
        for m in self.modules:
            m._grad_ready_hook = self._on_arena_ready
            if per_layer and hasattr(m, "_grad_layer_hook"):
                m._grad_layer_hook = self._on_layer_ready
        if self.bucket_blocks > 1:
            self._build_buckets()

    def _build_buckets(self) -> None:
        """Gated xattn blocks (modules without per-layer hand-over) in forward order, K per bucket: one flat fp32 buffer per bucket,
        the blocks' cached gradient arenas (FlatParams.grad_arena) become consecutive slices of it."""
        blocks = [m for m in self.modules if not hasattr(m, "_grad_layer_hook") and hasattr(getattr(m, "_fp", None), "total")]
        K = self.bucket_blocks
        for b0 in range(0, len(blocks), K):
            group = blocks[b0:b0 + K]
            flat = group[0]._fp.ensure()
            buf = torch.zeros(sum(m._fp.total for m in group), dtype=torch.float32, device=flat.device)
            off = 0
            for pos, m in enumerate(group):
                m._grad_arena = buf[off:off + m._fp.total]
                off += m._fp.total
                self._buckets[id(m)] = (len(self._bucket_state), pos)
            self._bucket_state.append({"buf": buf, "size": len(group), "ready": 0, "ptrs": [m._grad_arena.data_ptr() for m in group]})

    def detach(self) -> None:
        for m in self.modules:
            m._grad_ready_hook = None
            if hasattr(m, "_grad_layer_hook"):
                m._grad_layer_hook = None

    def no_sync(self):
        """Context manager, as DistributedDataParallel.no_sync(): backwards inside it exchange nothing (gradients accumulate
        locally in `.grad`); the first backward + finish() outside it averages the accumulated gradients once."""
        import contextlib

        @contextlib.contextmanager
        def ctx():
            old, self._sync = self._sync, False
            try:
                yield
            finally:
                self._sync = old
        return ctx()

    @staticmethod
    def _accumulating(module) -> bool:
        """True when this backward's gradients will be ADDED by autograd into existing `.grad`s (gradient accumulation,
        zero_grad(set_to_none=False)): the arena the backward just wrote is then a temporary that autograd reads right after
        the hook returns, so an asynchronous in-place all-reduce of it would race with that add.  Such modules are reduced in
        finish() instead, over the accumulated gradient - correct whether or not the earlier micro-batches were already
        averaged, because the mean over ranks of (a part common to all ranks + a local part) is the common part + the mean."""
        fp = getattr(module, "_fp", None)
        return hasattr(fp, "params") and any(p.grad is not None for p in fp.params())

    # called from inside the module's backward (autograd thread), right after its kernels were enqueued.
    # ranges: element ranges of the arena still to be reduced (None = all of it; per-layer callers pass what is left)
    def _on_arena_ready(self, module, arena: torch.Tensor, ranges=None) -> None:
        if self.world == 1 or not self._sync:
            return
        if self._accumulating(module):
            if module not in self._deferred:
                self._deferred.append(module)
            return
        slot = self._buckets.get(id(module)) if ranges is None else None
        if slot is not None:
            st = self._bucket_state[slot[0]]
            if arena.data_ptr() == st["ptrs"][slot[1]]:          # the backward wrote into its slice of the bucket
                st["ready"] += 1
                if st["ready"] == st["size"]:
                    st["ready"] = 0
                    self._launch(st["buf"])
                return
            # a fresh arena was allocated (gradient accumulation into existing .grad): reduce it on its own
        if ranges is None:
            self._launch(arena)
        else:
            for lo, hi in ranges:
                if hi > lo:
                    self._launch(arena[lo:hi])

    def _on_layer_ready(self, module, arena: torch.Tensor, lo: int, hi: int) -> None:
        if self.world > 1 and self._sync and hi > lo and not self._accumulating(module):
            self._launch(arena[lo:hi])

    def _launch(self, t: torch.Tensor) -> None:
        from . import functional as Fn
        if Fn._PENDING:                 # deferred side-stream join: the arena is complete only once the side stream is joined
            Fn.side_join()
        if t.is_cuda:
            if self.wire_dtype is not None and t.dtype == torch.float32 and self.wire_dtype != torch.float32:
                buf = t.to(self.wire_dtype)
                self.bytes_reduced += buf.numel() * buf.element_size()
                work = dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
                self._pending.append((work, None, (t, buf)))
                return
            self.bytes_reduced += t.numel() * t.element_size()
            work = dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
            self._pending.append((work, None, None))
        else:   # gloo has no AVG
            self.bytes_reduced += t.numel() * t.element_size()
            work = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._pending.append((work, t, None))

    def finish(self) -> None:
        """Call after loss.backward(): reduces the remaining (non-arena) trainable gradients, e.g. the token embedding
        the reference keeps trainable (modeling_flamingo.py:115), and waits for every outstanding collective."""
        from . import functional as Fn
        if Fn._PENDING:                 # a backward deferred its side-stream join
            Fn.side_join()
        if not self._sync:              # inside no_sync(): nothing was launched, nothing is exchanged
            return
        scatter = []
        if self.world > 1:
            for m in self._deferred:    # gradient accumulation: autograd has added this backward's part into .grad by now
                g, aliased = m._fp.current_grad(m)
                if g is not None:
                    self._launch(g)
                    if not aliased:
                        scatter.append((m, g))
            for p in self.extra_params:
                if p.grad is not None:
                    self._launch(p.grad)
        self._deferred.clear()
        for work, scale_me, restore in self._pending:
            work.wait()
            if scale_me is not None:
                scale_me.div_(self.world)
            if restore is not None:             # bf16 on the wire: the averaged values go back into the fp32 arena
                restore[0].copy_(restore[1])
        self._pending.clear()
        for m, g in scatter:
            m._fp.scatter_grad(g)
        if self._split is not None:
            self._split.finish()


class _LookupFn(torch.autograd.Function):
    """F.embedding whose weight gradient is NOT produced by autograd: backward hands (ids, d_rows) to the sink."""

    @staticmethod
    def forward(ctx, ids, weight, anchor, sink):
        ctx.sink = sink
        ctx.save_for_backward(ids)
        return torch.nn.functional.embedding(ids, weight)

    @staticmethod
    def backward(ctx, d_rows):
        (ids,) = ctx.saved_tensors
        ctx.sink._collect(ids, d_rows)
        return None, None, None, None


class SplitEmbeddingGrad:
    """Takes the tied token embedding (the one LM parameter the reference keeps trainable, modeling_flamingo.py:115) out
    of the gradient tail of a data-parallel step.

    The tied weight receives two gradient contributions: a DENSE one from the lm_head GEMM (final right at the START of
    backward) and a SPARSE one from the input lookup (B*S rows, final only at the very END of backward).  Left to
    autograd, their sum exists only when backward ends, so its (vocab x D) all-reduce cannot overlap anything.  Here
    the lookup is detached from the weight's autograd edge: the dense part is therefore complete early and is
    all-reduced while the rest of backward runs; the sparse part is exchanged as an all-gather of (ids, rows) —
    B*S*(D+1) elements per rank instead of vocab*D — and scatter-added locally.  The resulting weight.grad is the same
    mean-over-ranks gradient DistributedDataParallel would produce (up to summation order).

    Usage (bench.py / a trainer):  split = SplitEmbeddingGrad.install(model, reducer); ... loss.backward();
    reducer.finish()  (finish() calls split.finish()).

    Requirement: every rank feeds the SAME number of tokens per step (the row exchange is a fixed-size all_gather): pad batches
    to a fixed length (`processor(..., padding="max_length")`) when using this with real captions; ragged per-rank batches
    need the dense path (leave SplitEmbeddingGrad uninstalled)."""

    @classmethod
    def install(cls, model: torch.nn.Module, reducer: "GradArenaReducer") -> "SplitEmbeddingGrad":
        """Wire a FlamingoModel (or FlamingoBaseModel): its input embedding becomes the split lookup."""
        base = getattr(model, "flamingo", model)
        emb = base.lm.get_input_embeddings()
        split = cls(emb.weight, reducer, padding_idx=emb.padding_idx)
        base.embed_lookup = split.lookup
        return split

    def __init__(self, weight: torch.nn.Parameter, reducer: "GradArenaReducer", padding_idx: Optional[int] = None):
        """padding_idx: nn.Embedding.padding_idx of the lookup (OPT: 1) — that row receives no lookup gradient."""
        self.weight = weight
        self.padding_idx = padding_idx
        self.reducer = reducer
        self.group = reducer.group
        self.world = reducer.world
        self._anchor = torch.zeros((), device=weight.device, requires_grad=True)
        self._sparse = []
        self._dense_launched = False
        reducer.extra_params = [p for p in reducer.extra_params if p is not weight]
        reducer._split = self
        self._hook = weight.register_post_accumulate_grad_hook(self._on_dense_ready)

    def lookup(self, ids: torch.Tensor) -> torch.Tensor:
        if self._anchor.device != self.weight.device:
            self._anchor = torch.zeros((), device=self.weight.device, requires_grad=True)
        return _LookupFn.apply(ids, self.weight.detach(), self._anchor, self)

    def _collect(self, ids, d_rows) -> None:
        ids, rows = ids.reshape(-1), d_rows.reshape(-1, d_rows.shape[-1])
        if self.padding_idx is not None:       # static shapes (CUDA-graph friendly): zero the rows instead of dropping them
            rows = rows.masked_fill((ids == self.padding_idx).unsqueeze(-1), 0)
        self._sparse.append((ids, rows))

    def _on_dense_ready(self, param) -> None:     # autograd thread, right after the lm_head weight gradient was accumulated
        if self.world > 1 and param.grad is not None and self.reducer._sync:      # (inside no_sync(): the next synced backward reduces the sum)
            self.reducer._launch(param.grad)
        self._dense_launched = True

    def finish(self) -> None:
        """Exchange the lookup rows and add them to the (already reduced) dense gradient. Called by reducer.finish()
        after every outstanding all-reduce has completed."""
        w = self.weight
        for ids, rows in self._sparse:
            if w.grad is None:
                w.grad = torch.zeros_like(w)
            if self.world > 1:
                all_ids = torch.empty((self.world * ids.numel(),), dtype=ids.dtype, device=ids.device)
                all_rows = torch.empty((self.world * rows.shape[0], rows.shape[1]), dtype=rows.dtype, device=rows.device)
                dist.all_gather_into_tensor(all_ids, ids.contiguous(), group=self.group)
                dist.all_gather_into_tensor(all_rows, rows.contiguous(), group=self.group)
                self.reducer.bytes_reduced += all_ids.numel() * all_ids.element_size() + all_rows.numel() * all_rows.element_size()
                w.grad.index_add_(0, all_ids, all_rows.to(w.grad.dtype), alpha=1.0 / self.world)
            else:
                w.grad.index_add_(0, ids, rows.to(w.grad.dtype))
        self._sparse.clear()
        self._dense_launched = False

    def remove(self) -> None:
        self._hook.remove()
        self.reducer._split = None


def shard_batch(global_batch: int, rank: int, world: int) -> range:
    """Contiguous shard of sample indices owned by ``rank`` (weak scaling keeps per-rank size fixed instead)."""
    per = (global_batch + world - 1) // world
    return range(min(rank * per, global_batch), min((rank + 1) * per, global_batch))
