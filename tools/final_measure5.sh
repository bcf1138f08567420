#!/bin/bash
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; tail -4 gpurun_out/t_all.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c3_final.json 2> gpurun_out/bench_c3_final.err
for c in C1 C2 C4 C5; do
  python bench.py --config $c --steps $([ $c = C2 ] && echo 20 || ([ $c = C1 ] && echo 16 || echo 2)) --warmup 3 > gpurun_out/bench_${c}_final.json 2> gpurun_out/bench_${c}_final.err
done
python tools/append_probe.py 2>&1 | tail -2
python tools/panel_ab.py 2>&1 | tail -1
python - <<'PY'
import json
for f in ("bench_c3_final", "bench_C1_final", "bench_C2_final", "bench_C4_final", "bench_C5_final"):
    try:
        j = json.load(open(f"gpurun_out/{f}.json")); r = j.get("roofline", {})
        print(f, j["value"], j["unit"], "e2e", j["e2e"]["value"], j.get("clocks"), "frac", r.get("frac"), r.get("frac_of_sustained_peak"), "cpu", j.get("cpu_baseline", {}).get("value"), j.get("argmax_check_fp64_full_pool"), j.get("append"))
    except Exception as e:
        print(f, "FAILED", e)
PY
