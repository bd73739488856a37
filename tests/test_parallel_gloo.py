"""world_size-2 gloo test of the gradient-arena reducer (the N>1 host logic), on CPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flamingo_mini_b200.parallel import GradArenaReducer, shard_batch


class _FakeModule(torch.nn.Module):
    """Stands in for a hot-path module: owns an arena and calls the hook from 'backward'."""

    def __init__(self, n):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(n))
        self._fp = object()
        self._grad_ready_hook = None

    def fake_backward(self, value):
        arena = torch.full((self.w.numel(),), float(value))
        self.w.grad = arena
        if self._grad_ready_hook is not None:
            self._grad_ready_hook(self, arena)


def _worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mods = [_FakeModule(1000), _FakeModule(17)]
        extra = torch.nn.Parameter(torch.zeros(5))
        red = GradArenaReducer(mods, extra_params=[extra])
        for i, m in enumerate(reversed(mods)):           # backward order: last module first
            m.fake_backward(rank + 1 + i)
        extra.grad = torch.full((5,), 10.0 * (rank + 1))
        red.finish()
        # mean over ranks of (rank+1+i) = 1.5 + i
        assert torch.allclose(mods[1].w.grad, torch.full((17,), 1.5))
        assert torch.allclose(mods[0].w.grad, torch.full((1000,), 2.5))
        assert torch.allclose(extra.grad, torch.full((5,), 15.0))
        assert red.bytes_reduced == 4 * (1000 + 17 + 5)
        # a second step reuses the reducer
        for m in mods:
            m.fake_backward(4.0 if rank == 0 else 8.0)
        red.finish()
        assert torch.allclose(mods[0].w.grad, torch.full((1000,), 6.0))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_grad_arena_reducer_world2():
    mp.spawn(_worker, args=(2, _free_port()), nprocs=2, join=True)


def test_shard_batch():
    assert list(shard_batch(10, 0, 4)) == [0, 1, 2]
    assert list(shard_batch(10, 3, 4)) == [9]
    assert sum(len(shard_batch(32, r, 8)) for r in range(8)) == 32


# ---- split exchange of the tied token-embedding gradient (dense lm_head part early, lookup part as gathered rows)
class _TiedLM(torch.nn.Module):
    def __init__(self, vocab=37, d=8):
        super().__init__()
        torch.manual_seed(0)
        self.wte = torch.nn.Embedding(vocab, d)
        self.mix = torch.nn.Linear(d, d)
        self.embed_lookup = None

    def forward(self, ids):
        x = self.embed_lookup(ids) if self.embed_lookup is not None else self.wte(ids)
        h = torch.tanh(self.mix(x))
        return torch.nn.functional.linear(h, self.wte.weight)       # tied head


def _tied_loss(model, ids):
    logits = model(ids)
    return torch.nn.functional.cross_entropy(logits[:, :-1].reshape(-1, logits.size(-1)), ids[:, 1:].reshape(-1))


def _split_worker(rank, world, port):
    from flamingo_mini_b200.parallel import SplitEmbeddingGrad
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        all_ids = torch.randint(0, 37, (world, 3, 11), generator=g)        # every rank knows every shard (for the check)
        # reference: the mean over ranks of the plain single-process gradients
        ref_model = _TiedLM()
        want = torch.zeros_like(ref_model.wte.weight)
        for r in range(world):
            ref_model.zero_grad()
            _tied_loss(ref_model, all_ids[r]).backward()
            want += ref_model.wte.weight.grad / world
        # split exchange
        model = _TiedLM()
        for p in model.mix.parameters():
            p.requires_grad = False                                       # frozen LM body
        red = GradArenaReducer([], extra_params=[model.wte.weight])
        split = SplitEmbeddingGrad(model.wte.weight, red)
        assert red.extra_params == []                                     # the embedding left the dense tail
        model.embed_lookup = split.lookup
        for _ in range(2):                                                # second step reuses hook + sink
            model.zero_grad()
            _tied_loss(model, all_ids[rank]).backward()
            assert split._dense_launched
            red.finish()
            torch.testing.assert_close(model.wte.weight.grad, want, rtol=1e-5, atol=1e-6)
        # far fewer bytes than the dense (vocab x d) tail it replaces: here the dense part still travels (early, overlapped)
        assert red.bytes_reduced > 0
        # world-size-1 semantics (no process group needed for the math): same gradient as plain autograd
        split.remove()
    finally:
        dist.destroy_process_group()


def test_split_embedding_grad_world2():
    mp.spawn(_split_worker, args=(2, _free_port()), nprocs=2, join=True)


def test_split_embedding_grad_single_process():
    from flamingo_mini_b200.parallel import SplitEmbeddingGrad
    ids = torch.randint(0, 37, (2, 9), generator=torch.Generator().manual_seed(1))
    ref = _TiedLM(); _tied_loss(ref, ids).backward()
    model = _TiedLM()
    red = GradArenaReducer([], extra_params=[model.wte.weight])
    split = SplitEmbeddingGrad(model.wte.weight, red)
    model.embed_lookup = split.lookup
    _tied_loss(model, ids).backward()
    red.finish()
    torch.testing.assert_close(model.wte.weight.grad, ref.wte.weight.grad, rtol=1e-5, atol=1e-6)


# ---- per-layer hand-over of an arena (PerceiverResampler with fm_resampler_bwd_notify): layer slices first, rest at the end
class _LayeredFake(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(50))
        self._fp = object()
        self._grad_ready_hook = None
        self._grad_layer_hook = None
        self._layer_ranges = [(10, 20), (20, 30), (30, 40)]

    def fake_backward(self, rank):
        arena = torch.arange(50, dtype=torch.float32) * (rank + 1)
        self.w.grad = arena
        for lo, hi in reversed(self._layer_ranges):
            if self._grad_layer_hook is not None:
                self._grad_layer_hook(self, arena, lo, hi)
        if self._grad_ready_hook is not None:
            if self._grad_layer_hook is not None:
                self._grad_ready_hook(self, arena, [(0, 10), (40, 50)])
            else:
                self._grad_ready_hook(self, arena)


def _layered_worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for per_layer in (True, False):
            m = _LayeredFake()
            red = GradArenaReducer([m], per_layer=per_layer)
            assert (m._grad_layer_hook is not None) == per_layer
            m.fake_backward(rank)
            assert len(red._pending) == (5 if per_layer else 1)
            red.finish()
            torch.testing.assert_close(m.w.grad, torch.arange(50, dtype=torch.float32) * 1.5)     # mean of x1 and x2
            assert red.bytes_reduced == 200
            red.detach()
            assert m._grad_layer_hook is None and m._grad_ready_hook is None
    finally:
        dist.destroy_process_group()


def test_per_layer_arena_reduction_world2():
    mp.spawn(_layered_worker, args=(2, _free_port()), nprocs=2, join=True)


# ---- bucketed exchange: K consecutive blocks share one buffer and one all-reduce
def _bucket_worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from flamingo_mini_b200 import GatedCrossAttentionBlock
        torch.manual_seed(0)
        blocks = [GatedCrossAttentionBlock(dim=64, dim_visual=64) for _ in range(5)]
        red = GradArenaReducer(blocks, bucket_blocks=2)
        assert len(red._bucket_state) == 3 and [st["size"] for st in red._bucket_state] == [2, 2, 1]
        for step in range(2):
            for i in reversed(range(5)):                      # backward order: last block first
                m = blocks[i]
                arena = m._fp.grad_arena(m)                   # its slice of the bucket buffer
                assert arena.data_ptr() == red._bucket_state[i // 2]["ptrs"][i % 2]
                arena.fill_(float((rank + 1) * (i + 1) + step))
                launched = len(red._pending)
                m._grad_ready_hook(m, arena)
                # a bucket is reduced when its LAST block in backward order (= its first in forward order) is complete
                assert len(red._pending) == launched + (1 if i % 2 == 0 else 0)
            red.finish()
            for i, m in enumerate(blocks):
                want = 1.5 * (i + 1) + step                    # mean over ranks 0, 1
                assert torch.allclose(m._fp.grad_arena(m), torch.full((m._fp.total,), want)), (i, step)
    finally:
        dist.destroy_process_group()


def test_bucketed_arena_reduce_world2():
    mp.spawn(_bucket_worker, args=(2, _free_port()), nprocs=2, join=True)


# ---- gradient accumulation: a backward that adds into existing .grad is reduced in finish(), over the accumulated gradient
def _emulate_backward(m, fill):
    """What functional._XattnFn.backward + autograd's AccumulateGrad do, without the CUDA library: the backward writes an arena
    (the module's cached one when no .grad exists yet, a fresh one otherwise), hands it to the reducer's hook, returns views of
    it, and autograd either installs those views as .grad or adds them into the existing .grad."""
    g = m._fp.grad_arena(m)
    g.fill_(float(fill))
    m._last_grad_arena = g
    if m._grad_ready_hook is not None:
        m._grad_ready_hook(m, g)
    for p, v in zip(m._fp.params(), m._fp.grad_views(g)):
        if p.grad is None:
            p.grad = v
        else:
            p.grad.add_(v)
    return g


def _accum_worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from flamingo_mini_b200 import GatedCrossAttentionBlock
        torch.manual_seed(0)
        a, b = (1.0, 3.0)[rank], (10.0, 30.0)[rank]              # micro-batch gradients of this rank; means over ranks: 2 and 20
        # (i) finish() after every backward; (ii) first micro-batch inside no_sync(); (iii) .grad pre-assigned by the caller
        for mode in ("finish_each", "no_sync", "foreign_grads"):
            m = GatedCrossAttentionBlock(dim=64, dim_visual=64)
            red = GradArenaReducer([m])
            if mode == "foreign_grads":                          # zero_grad(set_to_none=False) on gradients that are not arena views
                for p in m.parameters():
                    p.grad = torch.zeros_like(p)
            if mode == "no_sync":
                with red.no_sync():
                    _emulate_backward(m, a)
                    red.finish()
                assert not red._pending and red.bytes_reduced == 0
            else:
                _emulate_backward(m, a)
                assert len(red._pending) == (0 if mode == "foreign_grads" else 1)
                red.finish()
            g2 = _emulate_backward(m, b)
            assert g2 is not getattr(m, "_grad_arena", None)     # accumulation: a temporary arena, NOT reduced in place
            assert not red._pending and red._deferred == [m]
            red.finish()
            assert not red._deferred
            for n, p in m.named_parameters():
                assert torch.allclose(p.grad, torch.full_like(p, 22.0)), (mode, n, p.grad.flatten()[:3])
            assert torch.allclose(g2, torch.full_like(g2, b))    # the temporary was left alone
            flat, aliased = m._fp.current_grad(m)
            assert aliased == (mode != "foreign_grads") and torch.allclose(flat[:m._fp.slots[-1][1]], torch.full((m._fp.slots[-1][1],), 22.0))
    finally:
        dist.destroy_process_group()


def test_gradient_accumulation_world2():
    mp.spawn(_accum_worker, args=(2, _free_port()), nprocs=2, join=True)


# ---- the whole N > 1 host path with the REAL modules: the body of tests/test_gpu_nccl.py on CPU tensors, the CUDA library
# replaced by its host emulation (tests/_emu_util.py) and NCCL by gloo.  Covers what the fakes above cannot: hooks fired from
# inside the real autograd backward, per-layer hand-over through the fm_resampler_bwd_notify C callback, the split embedding
# exchange inside FlamingoModel.forward.
def _dp_emu_step(model, clip, ids, ml, reducer=None, micro_batches=1):
    import tests.test_gpu_nccl as T
    model.zero_grad(set_to_none=True)
    n = ids.shape[0] // micro_batches
    for i in range(micro_batches):                                    # > 1: gradient accumulation, finish() after every backward
        sl, csl = slice(i * n, (i + 1) * n), slice(i * n * T.TINY["N"], (i + 1) * n * T.TINY["N"])
        vf = model.flamingo.resampler(clip[csl]).reshape(n, T.TINY["N"], 64, T.TINY["Dv"])
        out = model(input_ids=ids[sl], media_locations=ml[sl], visual_features=vf, labels=ids[sl], attention_mask=torch.ones_like(ids[sl]))
        (out.loss / micro_batches).backward()
        if reducer is not None:
            reducer.finish()
    return {n_: p.grad.detach().float().clone() for n_, p in model.named_parameters() if p.requires_grad and p.grad is not None}


def _dp_emu_worker(rank, world, port, per_layer, split, micro_batches, out_path):
    import tests.test_gpu_nccl as T
    from tests import _emu_util
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        with _emu_util.swapped_in():
            from flamingo_mini_b200.parallel import SplitEmbeddingGrad, hot_path_modules
            model = T._model("cpu")
            hot = hot_path_modules(model)
            hot_ids = {id(p) for m in hot for p in m.parameters()}
            extra = [p for p in model.parameters() if p.requires_grad and id(p) not in hot_ids]
            red = GradArenaReducer(hot, extra_params=extra, per_layer=per_layer)
            if split:
                SplitEmbeddingGrad.install(model, red)
            clip, ids, ml = T._batch(2 * world)
            sl = slice(rank * 2, rank * 2 + 2)
            csl = slice(rank * 2 * T.TINY["N"], (rank * 2 + 2) * T.TINY["N"])
            got = _dp_emu_step(model, clip[csl], ids[sl], ml[sl], red, micro_batches)
            assert red.bytes_reduced > 0 and not red._pending and not red._deferred
            if rank == 0:
                torch.save(got, out_path)
            dist.barrier()
    finally:
        dist.destroy_process_group()


import pytest  # noqa: E402

_DP_EMU_CASES = [(True, True, 1)] + ([(False, False, 1), (True, False, 2)] if os.environ.get("FM_EMU_SLOW") else [])


@pytest.mark.parametrize("per_layer,split,micro_batches", _DP_EMU_CASES)
def test_data_parallel_step_equals_single_process_through_the_emulator(tmp_path, per_layer, split, micro_batches):
    from tests import _emu_util
    if not _emu_util.available():
        pytest.skip("g++ / CUDA headers not available")
    import tests.test_gpu_nccl as T
    _emu_util.build()                                                 # once, before the ranks race for it
    world, out_path = 2, str(tmp_path / "rank0_grads.pt")
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_dp_emu_worker, args=(r, world, port, per_layer, split, micro_batches, out_path)) for r in range(world)]
    for p in procs:
        p.start()
    with _emu_util.swapped_in():                                      # meanwhile: one process, the concatenated batch, same kernels
        model = T._model("cpu")
        clip, ids, ml = T._batch(2 * world)
        ref = _dp_emu_step(model, clip, ids, ml)
    for p in procs:
        p.join(timeout=900)
        assert p.exitcode == 0
    got = torch.load(out_path)
    assert set(got) == set(ref) and len(ref) > 40
    for n in ref:
        d = (got[n] - ref[n]).norm().item() / (ref[n].norm().item() + 1e-12)
        assert d < 2e-3, f"{n}: rank-averaged gradient differs from the single-process gradient by {d:.3e}"
