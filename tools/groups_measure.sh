#!/bin/bash
for G in 4 8 2; do
  BO_B200_LML_GROUPS=$G python bench.py --config C5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C5_g$G.json 2> gpurun_out/bench_C5_g$G.err
  python -c "
import json; j=json.load(open('gpurun_out/bench_C5_g$G.json')); print('C5 groups=$G', j['value'], j['roofline']['frac'])"
done
for G in 4 8; do BO_B200_LML_GROUPS=$G python tools/panel_ab.py 2>&1 | tail -1; done
