"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares of device time)."""
import csv, collections, io, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(io.StringIO("".join(lines))):
    name = row["Kernel Name"].split("(")[0][:80]
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(u, 1.0)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"total device time {tot:.3f} ms over {sum(a[0] for a in agg.values())} launches (ncu: cold-cache, serialised -- compare shares)")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v:14.3f} ms {100 * v / tot:7.3f}%  x{c:5d}  {k}")
