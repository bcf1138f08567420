#!/bin/bash
BO_B200_CHOL_LEFT=0 python tools/panel_ab.py 2>&1 | tail -1
BO_B200_CHOL_LEFT=1 python tools/panel_ab.py 2>&1 | tail -1
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_extras.py tests/test_gpu_n4.py tests/test_gpu_fuzz.py tests/test_gpu_optimizer.py tests/test_gpu_svgp.py -q -x 2>&1 | tail -8
for c in C1 C5; do
  python bench.py --config $c --steps $([ $c = C1 ] && echo 8 || echo 2) --warmup 3 --no-cpu-baseline > gpurun_out/bench_${c}_left.json 2> gpurun_out/bench_${c}_left.err
  python -c "
import json; j=json.load(open('gpurun_out/bench_${c}_left.json')); print('$c', j['value'], j['roofline']['frac'], j.get('refit_ms'))"
done
