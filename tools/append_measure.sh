#!/bin/bash
python tools/append_probe.py 2>&1 | tail -2
timeout 1200 python -m pytest tests/test_gpu_extras.py tests/test_gpu_n4.py tests/test_gpu_optimizer.py tests/test_gpu_fuzz.py tests/test_gpu_parity.py -q -x 2>&1 | tail -4
python bench.py --config C4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C4_final.json 2> gpurun_out/bench_C4_final.err
python -c "
import json; j=json.load(open('gpurun_out/bench_C4_final.json')); print('C4', j['value'], j['e2e']['value'], j['append'])"
