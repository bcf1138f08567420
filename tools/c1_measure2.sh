#!/bin/bash
python tools/append_probe.py 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_optimizer.py tests/test_gpu_extras.py -q -x 2>&1 | tail -3
python bench.py --config C1 --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C1.json 2> gpurun_out/bench_C1.err; tail -3 gpurun_out/bench_C1.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_C1.json"))
print(j["value"], j["unit"], "e2e", j["e2e"]["value"], "roof", j["roofline"]["frac"], j["roofline"]["evaluations_per_step"], j["roofline"]["share_of_step"])
print(json.dumps(j["hyperfit"], indent=1)); print(json.dumps(j["phases"], indent=1))
PY
