#!/bin/bash
for F in 2 74 2 74; do
  BO_B200_SWEEP_FLAGS=$F python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_hint$F.json 2> gpurun_out/bench_c3_hint$F.err
  python -c "
import json; j=json.load(open('gpurun_out/bench_c3_hint$F.json')); r=j['roofline']; print('flags=$F', j['value'], j['clocks']['sm_mhz'], r['frac'], r.get('frac_of_sustained_peak'), j['roofline'].get('kernel_ms'))"
done
