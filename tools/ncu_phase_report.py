"""Summarise an ncu report of sweep_kernel: pipe utilisation + warp-sample share per phase (development aid)."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for h, u, v in zip(hdr, units, vals):
    if h in want: print(f"{h:85s} {u:10s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
iS, iN = hdr.index("Source"), hdr.index("# Samples")
tot = sum(int(r[iN]) for r in data)
first_ublk = min(k for k, r in enumerate(data) if "UBLKCP" in r[iS])
dm = [k for k, r in enumerate(data) if "DMMA" in r[iS]]
def s(a, b): return 100.0 * sum(int(r[iN]) for r in data[a:b]) / tot
print(f"samples: before first UBLKCP (setup + phase A) {s(0, first_ublk):.2f}% | UBLKCP..last DMMA (phase B) {s(first_ublk, dm[-1]+1):.2f}% | after (reduce/epilogue/topk) {s(dm[-1]+1, len(data)):.2f}%")
stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
for name, (a, b) in {"phaseA": (0, first_ublk), "phaseB": (first_ublk, dm[-1] + 1), "after": (dm[-1] + 1, len(data))}.items():
    st = collections.Counter()
    for r in data[a:b]:
        for c in stall_cols:
            v = r[hdr.index(c)]
            if v and v != "0": st[c] += int(v)
    print(name, [(k, round(100.0 * v / tot, 2)) for k, v in st.most_common(5)])
