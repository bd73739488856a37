"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals for ONE step.
A step starts at fm::bcast_rows_kernel (first kernel of the resampler forward); the last complete step is used.
usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt"""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    starts = [i for i, n in enumerate(names) if "bcast_rows_kernel" in n]
    if len(starts) >= 2:
        lo, hi = starts[-2], starts[-1]
    else:
        lo, hi = 0, len(rows)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[lo:hi]:
        v = float(r["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "usecond": v, "ms": v * 1e3}.get(r["Metric Unit"], v)
        short = re.sub(r"\(.*", "", r["Kernel Name"])
        short = re.sub(r"^void ", "", short)[:90]
        agg[short][0] += 1
        agg[short][1] += v
    tot = sum(v[1] for v in agg.values())
    fm = sum(v[1] for k, v in agg.items() if k.startswith("fm::"))
    print(f"launches in step: {hi - lo}; total kernel time {tot/1e3:.3f} ms; fm:: kernels {fm/1e3:.3f} ms ({100*fm/tot:.1f} %)")
    print("(ncu per-launch times are cold-cache and serialised: compare SHARES with bench.py's in-situ numbers)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"{v[1]:10.1f} us {100*v[1]/tot:5.1f}% n={v[0]:4d} {k}")


if __name__ == "__main__":
    main(sys.argv[1])
