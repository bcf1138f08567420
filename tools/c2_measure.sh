#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_i8.py tests/test_gpu_i8_refdata.py tests/test_gpu_fuzz.py -q -x 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --config C2 --steps 20 --warmup 5 > gpurun_out/bench_C2_final.json 2> gpurun_out/bench_C2_final.err
python -c "
import json; j=json.load(open('gpurun_out/bench_C2_final.json')); print('C2', j['value'], 'e2e', j['e2e']['value'], j['roofline']['frac'], j.get('argmax_check_fp64_full_pool'))"
