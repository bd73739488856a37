#!/usr/bin/env python
"""Pick the roofline-relevant metrics out of `ncu -i X.ncu-rep --page raw --csv` (stdin) and print one line per kernel launch.

    ncu -i gpurun_out/ncu/act_gemm.ncu-rep --page raw --csv | python tools/ncu_pick.py
"""
import csv
import sys

WANT = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg")


def main():
    rows = list(csv.reader(sys.stdin))
    if len(rows) < 3:
        print("no data")
        return
    head, units = rows[0], rows[1]
    cols = [i for i, h in enumerate(head) if any(h.startswith(w) for w in WANT)]
    name_col = head.index("Kernel Name") if "Kernel Name" in head else None
    for r in rows[2:]:
        if len(r) != len(head):
            continue
        print(r[name_col][:90] if name_col is not None else "kernel")
        for i in cols:
            print(f"    {head[i]:72s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    main()
