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
