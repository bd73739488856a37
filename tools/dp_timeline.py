#!/usr/bin/env python
"""Where does a data-parallel step lose time?  Run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 tools/dp_timeline.py

Captures the C2 training step (default gradient exchange) as a CUDA graph exactly like bench.py, replays it under CUPTI activity
tracing on every rank, and rank 0 prints: the step's span, the NCCL kernels (count, summed duration, union of their busy
intervals), how much of that union lies AFTER the last compute kernel has ended (the exposed tail), and the five longest collectives.
"""
from __future__ import annotations

import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def union(intervals):
    intervals = sorted(intervals)
    total, cur_s, cur_e = 0.0, None, None
    for s, e in intervals:
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                total += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    if cur_e is not None:
        total += cur_e - cur_s
    return total


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from flamingo_mini_b200.parallel import GradArenaReducer, SplitEmbeddingGrad, hot_path_modules
    w = bench.WORKLOADS[os.environ.get("WORKLOAD", "c2")]
    model = bench.build_model(w, dev)
    hot = hot_path_modules(model)
    hot_ids = {id(p) for m in hot for p in m.parameters()}
    extra = [p for p in model.parameters() if p.requires_grad and id(p) not in hot_ids]
    reducer = None
    if world > 1:
        reducer = GradArenaReducer(hot, extra_params=extra, per_layer=True, bucket_blocks=int(os.environ.get("BUCKET", "1")))
        SplitEmbeddingGrad.install(model, reducer)
    clip, ids, ml = bench.make_batch(w, w["B"], dev, 1234 + rank, torch.bfloat16)
    for m in hot:
        m._fp.always_refresh = True

    def step():
        model.zero_grad(set_to_none=True)
        return bench.train_step(model, w, clip, ids, ml, reducer)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    model.zero_grad(set_to_none=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        bench.train_step(model, w, clip, ids, ml, reducer)
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        g.replay()
        g.replay()
        torch.cuda.synchronize()
    evs = [(e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.time_range.end > e.time_range.start and "Memcpy" not in e.name and "Memset" not in e.name]
    evs.sort()
    # second replay only: everything that starts after the midpoint between the two replays' spans
    starts = [s for s, _, _ in evs]
    mid = (starts[0] + max(e for _, e, _ in evs)) / 2
    evs = [x for x in evs if x[0] >= mid]
    nccl = [(s, e, n) for s, e, n in evs if "nccl" in n.lower()]
    comp = [(s, e, n) for s, e, n in evs if "nccl" not in n.lower()]
    t0, t1 = min(s for s, _, _ in evs), max(e for _, e, _ in evs)
    last_comp = max(e for _, e, _ in comp)
    tail = union([(max(s, last_comp), e) for s, e, _ in nccl if e > last_comp])
    out = {"world": world, "span_us": t1 - t0, "compute_kernels": len(comp), "compute_sum_us": sum(e - s for s, e, _ in comp),
           "compute_union_us": union([(s, e) for s, e, _ in comp]),
           "nccl_kernels": len(nccl), "nccl_sum_us": sum(e - s for s, e, _ in nccl), "nccl_union_us": union([(s, e) for s, e, _ in nccl]),
           "nccl_after_last_compute_us": tail,
           "longest_nccl_us": sorted(((e - s), n[:60]) for s, e, n in nccl)[-5:][::-1]}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        g = None
        del model, reducer
        import gc
        gc.collect()
        torch.cuda.synchronize()
        import threading
        wd = threading.Timer(30.0, lambda: os._exit(0))
        wd.daemon = True
        wd.start()
        dist.destroy_process_group()
        wd.cancel()


if __name__ == "__main__":
    main()
