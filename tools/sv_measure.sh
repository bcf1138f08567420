#!/bin/bash
# SVGP sliced sweep bring-up: parity tests, then the sliced regressions and a short C3 line (the pair kernel changed)
export BO_I8_WAIT_CYCLES=4000000000
timeout 900 python -m pytest tests/test_gpu_svgp.py -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_i8.py tests/test_gpu_i8_refdata.py tests/test_gpu_n4.py -x -q 2>&1 | tail -5
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_sv.json 2> gpurun_out/bench_c3_sv.err
python -c "
import json; j=json.load(open('gpurun_out/bench_c3_sv.json')); print('C3', j['value'], j['clocks']['sm_mhz'], j['roofline']['frac'])"
python tools/svgp_sweep_timing.py 2>&1 | tail -12
