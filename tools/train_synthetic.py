#!/usr/bin/env python
"""Synthetic-data training run of the hot path (the reference's training/train.sh recipe without COCO):

    python tools/train_synthetic.py --workload c2 --steps 20
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_synthetic.py ...

forward -> loss -> backward -> gradient all-reduce (N > 1) -> ArenaAdamW step, on bench.py's synthetic batches
(random CLIP patch features, random token ids).  Prints the loss curve and samples/s INCLUDING the optimizer step;
the benchmark metric (fwd+bwd only) is bench.py's."""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from flamingo_mini_b200.training import save_trainable, train  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2", choices=sorted(bench.WORKLOADS))
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--lr", type=float, default=5e-5)
    ap.add_argument("--warmup-steps", type=int, default=0)
    ap.add_argument("--split-embedding", action="store_true")
    ap.add_argument("--save", default=None, help="write a trainable-only checkpoint here when done")
    args = ap.parse_args()
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w = bench.WORKLOADS[args.workload]
    model = bench.build_model(w, dev)
    clip, ids, ml = bench.make_batch(w, w["B"], dev, 1234 + rank, torch.bfloat16)

    def batches():
        while True:
            vf = model.flamingo.resampler(clip).reshape(ids.shape[0], w["N"], 64, w["Dv"])
            yield dict(input_ids=ids, media_locations=ml, visual_features=vf, labels=ids, attention_mask=torch.ones_like(ids))

    torch.cuda.synchronize()
    t0 = time.time()
    losses = train(model, batches(), steps=args.steps, lr=args.lr, warmup_steps=args.warmup_steps, log_every=max(1, args.steps // 10),
                   split_embedding=args.split_embedding)
    torch.cuda.synchronize()
    dt = time.time() - t0
    if rank == 0:
        print(f"{args.steps} steps in {dt:.2f} s: {w['B'] * world * args.steps / dt:.1f} samples/s incl. optimizer; "
              f"loss {losses[0]:.4f} -> {losses[-1]:.4f}")
        if args.save:
            save_trainable(model, args.save, step=args.steps)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
