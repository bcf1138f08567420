#!/bin/bash
# C1 bench line (BO iteration on the reference's CSV rows) + its reference arm
python bench.py --config C1 --steps 5 --warmup 3 > gpurun_out/bench_C1.json 2> gpurun_out/bench_C1.err; tail -3 gpurun_out/bench_C1.err
python bench.py --config C1 --impl reference --steps 1 --warmup 0 > gpurun_out/bench_C1_reference.json 2> gpurun_out/bench_C1_reference.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_C1.json"))
print(j["value"], j["unit"], "e2e", j["e2e"]["value"], "roof", j["roofline"]["frac"], j["roofline"]["evaluations_per_step"], j["roofline"]["share_of_step"])
print(json.dumps(j["hyperfit"], indent=1)); print(json.dumps(j["phases"], indent=1)); print(j.get("cpu_baseline"))
PY
