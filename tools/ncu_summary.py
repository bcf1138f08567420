#!/usr/bin/env python
"""Side-by-side summary of `ncu --page raw --csv` exports (one kernel each).

usage: python tools/ncu_summary.py label=profiles/a_raw.csv label2=profiles/b_raw.csv [--candidates 18944]
Prints the rows DESIGN.md quotes (duration, cycles, tensor/FP64 pipe activity, DRAM bytes,
L2 hit rate, shared-memory wavefronts, instructions, launch shape) and the DRAM bytes per candidate.
"""
import csv
import sys

ROWS = [
    ("kernel duration", "gpu__time_duration.sum"),
    ("SM cycles elapsed (max)", "sm__cycles_elapsed.max"),
    ("SM clock", "sm__cycles_elapsed.max.per_second"),
    ("tensor pipe active (% of elapsed)", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("tensor memory path active (% of elapsed)", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("FP64 pipe active (% of active)", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("DRAM read", "dram__bytes_read.sum"),
    ("DRAM written", "dram__bytes_write.sum"),
    ("DRAM throughput (% of peak)", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 hit rate", "lts__t_sector_hit_rate.pct"),
    ("L2 throughput (% of peak)", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("shared-memory wavefronts LSU", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    ("shared-memory bank conflicts LSU", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("warp instructions executed", "smsp__inst_executed.sum"),
    ("registers per thread (launch)", "launch__registers_per_thread"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("cluster size", "launch__cluster_size"),
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    vals = rows[-1]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def main():
    cand = None
    cols = []
    args = sys.argv[1:]
    while args:
        a = args.pop(0)
        if a == "--candidates":
            cand = int(args.pop(0))
        else:
            label, path = a.split("=", 1)
            cols.append((label, path, load(path)))
    for label, path, m in cols:
        print(f"{label}: {m['Kernel Name'][0]}  <- {path}")
    print()
    for name, key in ROWS:
        cells = []
        for _, _, m in cols:
            v, u = m.get(key, ("n/a", ""))
            cells.append(f"{v} {u}".strip())
        print(f"{name:<44}" + " | ".join(f"{c:>24}" for c in cells))
    if cand:
        print()
        for label, _, m in cols:
            rd = float(m["dram__bytes_read.sum"][0]) * SCALE[m["dram__bytes_read.sum"][1]]
            wr = float(m["dram__bytes_write.sum"][0]) * SCALE[m["dram__bytes_write.sum"][1]]
            ms = float(m["gpu__time_duration.sum"][0])
            print(f"{label}: DRAM per candidate {(rd + wr) / cand:.0f} B (read {rd / cand:.0f} + written {wr / cand:.0f}); "
                  f"{cand / ms / 1e3:.3f} M candidates/s under ncu at the captured clock")


if __name__ == "__main__":
    main()
