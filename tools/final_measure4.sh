#!/bin/bash
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; tail -4 gpurun_out/t_all.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c3_final.json 2> gpurun_out/bench_c3_final.err
for c in C1 C2 C4 C5; do
  python bench.py --config $c --steps $([ $c = C2 ] && echo 20 || ([ $c = C1 ] && echo 8 || echo 2)) --warmup 3 > gpurun_out/bench_${c}_final.json 2> gpurun_out/bench_${c}_final.err
done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c3_reference.json 2> gpurun_out/bench_c3_reference.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sweep|topk|i8_" -c 200 --csv --log-file gpurun_out/launches_bench_sweeps.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
export PROF_POOL=18944
python tools/profile_sweep.py > gpurun_out/plain_pair.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sweep_i8 -s 1 -c 1 -f -o /tmp/prof_i8_pair python tools/profile_sweep.py > gpurun_out/ncu_pair.log 2>&1
ncu -i /tmp/prof_i8_pair.ncu-rep --page raw --csv > gpurun_out/ncu_i8_pair7_raw.csv 2>/dev/null
ncu -i /tmp/prof_i8_pair.ncu-rep --page details > gpurun_out/ncu_i8_pair7_details.txt 2>/dev/null
python - <<'PY'
import json
for f in ("bench_c3_final", "bench_C1_final", "bench_C2_final", "bench_C4_final", "bench_C5_final", "bench_c3_reference"):
    try:
        j = json.load(open(f"gpurun_out/{f}.json")); r = j.get("roofline", {})
        print(f, j["value"], j["unit"], "e2e", j["e2e"]["value"], j.get("clocks"), "frac", r.get("frac"), r.get("frac_of_sustained_peak"), "traffic", r.get("traffic"), "cpu", j.get("cpu_baseline", {}).get("value"), j.get("argmax_check_fp64_full_pool"))
    except Exception as e:
        print(f, "FAILED", e)
PY
