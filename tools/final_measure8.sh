#!/bin/bash
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c3_final.json 2> gpurun_out/bench_c3_final.err
python bench.py --config C4 --steps 2 --warmup 3 > gpurun_out/bench_C4_final.json 2> gpurun_out/bench_C4_final.err
timeout 300 python tools/svgp_scan_timing.py 2>&1 | tail -9
export PROF_POOL=18944
python tools/profile_sweep.py > gpurun_out/plain_pair.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sweep_i8 -s 1 -c 1 -f -o /tmp/prof_i8_pair python tools/profile_sweep.py > gpurun_out/ncu_pair.log 2>&1
ncu -i /tmp/prof_i8_pair.ncu-rep --page raw --csv > gpurun_out/ncu_i8_pair7_raw.csv 2>/dev/null
ncu -i /tmp/prof_i8_pair.ncu-rep --page details > gpurun_out/ncu_i8_pair7_details.txt 2>/dev/null
python - <<'PY'
import json
for f in ("bench_c3_final", "bench_C4_final"):
    j = json.load(open(f"gpurun_out/{f}.json")); r = j.get("roofline", {})
    print(f, j["value"], j["unit"], "e2e", j["e2e"]["value"], j.get("clocks"), "frac", r.get("frac"), r.get("frac_of_sustained_peak"), "cpu", j.get("cpu_baseline", {}).get("value"), j.get("argmax_check_fp64_full_pool"))
PY
