#!/bin/bash
python bench.py --config C5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C5_plain.json 2> gpurun_out/bench_C5_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_c5.csv python bench.py --config C5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c5.log 2>&1
python - <<'PY'
import csv, collections, re, json
print(json.load(open("gpurun_out/bench_C5_plain.json"))["value"])
rows=[r for r in csv.reader(open("gpurun_out/launches_c5.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); gi=hdr.index("Grid Size")
# keep the launches of the LAST step only: split on rescale_batched_kernel occurrences
names=[re.sub(r"\(.*","",r[ki]) for r in rows[1:]]
agg=collections.OrderedDict(); tot=0
for r in rows[1:]:
    ms=float(r[vi].replace(",",""))/1e6; k=re.sub(r"\(.*","",r[ki])
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=ms; tot+=ms
print(f"total {tot:.2f} ms over {len(rows)-1} launches (4 steps: 3 warm-up + 1)")
for k,(n,ms) in sorted(agg.items(), key=lambda x:-x[1][1])[:14]:
    print(f"{ms:10.3f} ms {100*ms/tot:6.2f}%  x{n:5d}  avg {1e3*ms/n:8.1f} us  {k}")
PY
