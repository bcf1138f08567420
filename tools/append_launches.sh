#!/bin/bash
python tools/append_probe.py 2>&1 | tail -2 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"kq_build|trmv|append|pack|pad_identity" --csv --log-file gpurun_out/launches_append.csv python tools/append_probe.py > gpurun_out/ncu_append.log 2>&1
python - <<'PY'
import csv, collections, re
rows=[r for r in csv.reader(open("gpurun_out/launches_append.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); gi=hdr.index("Grid Size")
seq=[(re.sub(r"\(.*","",r[ki]).replace("void ",""), r[gi], float(r[vi])/1e3) for r in rows[1:]]
# print the launches of the last two appends at each n
names=[s[0] for s in seq]
idx=[i for i,nm in enumerate(names) if nm.startswith("kq_build")]
for a in (idx[18], idx[-2]):
    b=idx[idx.index(a)+1]
    print(" | ".join(f"{nm[:22]} {g} {us:.1f}us" for nm,g,us in seq[a:b]), " total %.1f us" % sum(us for _,_,us in seq[a:b]))
PY
