#!/bin/bash
# L2 eviction-priority hints on the pair kernel's TMA loads (flags: 2 = default; +8 / +16 = L^-1 evict_last / evict_first; +32 / +64 = panel)
for F in 2 10 34 74 26 42; do
  BO_B200_SWEEP_FLAGS=$F python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_hint$F.json 2> gpurun_out/bench_c3_hint$F.err
  python -c "
import json; j=json.load(open('gpurun_out/bench_c3_hint$F.json')); r=j['roofline']; print('flags=$F', j['value'], j['clocks']['sm_mhz'], r['frac'], r.get('frac_of_sustained_peak'))"
done
