"""One eager C2 step under torch.profiler: where does the non-library GPU time go (frozen LM, loss head, autograd glue)?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    w = bench.WORKLOADS[os.environ.get("WORKLOAD", "c2")]
    dev = torch.device("cuda", 0)
    model = bench.build_model(w, dev)
    if os.environ.get("SIDE_STREAM", "0") == "0":       # per-kernel durations are only clean when kernels do not overlap
        from flamingo_mini_b200 import _lib
        _lib.load().fm_set_option(0, 0)
    clip, ids, ml = bench.make_batch(w, w["B"], dev, 1234, torch.bfloat16)
    for _ in range(3):
        model.zero_grad(set_to_none=True)
        bench.train_step(model, w, clip, ids, ml)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        model.zero_grad(set_to_none=True)
        bench.train_step(model, w, clip, ids, ml)
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", 0) or getattr(e, "cuda_time_total", 0)
        if t > 0 and e.device_type.name == "CUDA":
            rows.append((t, e.count, e.key))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    fm = sum(r[0] for r in rows if "fm::" in r[2])
    print(f"total CUDA kernel time {tot/1e3:.2f} ms; fm:: kernels {fm/1e3:.2f} ms; other {(tot-fm)/1e3:.2f} ms")
    for t, n, k in rows[:45]:
        print(f"{t/1e3:8.3f} ms n={n:4d} {k[:110]}")


if __name__ == "__main__":
    main()
