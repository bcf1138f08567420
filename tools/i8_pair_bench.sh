#!/bin/bash
echo "--- correctness: small shapes, pair"; timeout 120 python tools/i8_sweep_check.py 1000 5 200000 8 2>&1 | tail -6
timeout 120 python tools/i8_sweep_check.py 300 3 777 7 2>&1 | tail -6
echo "--- correctness: one-CTA"; BO_B200_I8_PAIR=0 timeout 120 python tools/i8_sweep_check.py 1000 5 200000 8 2>&1 | tail -6
echo "--- C3 accounting pair / one"; BO_B200_SWEEP_FLAGS=6 timeout 300 python tools/i8_sweep_check.py 4096 8 2400000 8 2>&1 | grep -E "sweep_i8|i8: sweep"
BO_B200_I8_PAIR=0 BO_B200_SWEEP_FLAGS=6 timeout 300 python tools/i8_sweep_check.py 4096 8 2400000 8 2>&1 | grep -E "sweep_i8|i8: sweep"
echo "--- bench C3 pair"; python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pair.json 2> gpurun_out/bench_pair.err; python -c "
import json; j=json.load(open('gpurun_out/bench_pair.json')); print(j['value'], j['e2e']['value'], j['clocks'], j['roofline']['frac'], j['roofline']['frac_of_sustained_peak'], j['roofline']['peak'], j['roofline']['peak_sustained'], j['argmax_check_fp64_full_pool'])"
echo "--- bench C3 one-CTA"; BO_B200_I8_PAIR=0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_one.json 2> gpurun_out/bench_one.err; python -c "
import json; j=json.load(open('gpurun_out/bench_one.json')); print(j['value'], j['e2e']['value'], j['clocks'], j['roofline']['frac'], j['roofline']['frac_of_sustained_peak'], j['argmax_check_fp64_full_pool'])"
