#!/bin/bash
export BO_I8_WAIT_CYCLES=4000000000
timeout 900 python -m pytest tests/test_gpu_svgp.py -q 2>&1 | tail -5
timeout 300 python tools/svgp_scan_timing.py 2>&1 | tail -10
timeout 600 python -m pytest tests/test_gpu_extras.py tests/test_gpu_optimizer.py -q 2>&1 | tail -5
python bench.py --config C1 --steps 8 --warmup 3 > gpurun_out/bench_C1.json 2> gpurun_out/bench_C1.err; tail -3 gpurun_out/bench_C1.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_C1.json"))
print(j["value"], j["unit"], "e2e", j["e2e"]["value"], "roof", j["roofline"]["frac"], j["roofline"]["evaluations_per_step"], j["roofline"]["share_of_step"])
print(json.dumps(j["hyperfit"], indent=1)); print(json.dumps(j["phases"], indent=1)); print(j.get("cpu_baseline"))
PY
