#!/bin/bash
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_final2.json 2> gpurun_out/bench_c3_final2.err
for c in C1 C2 C5; do
  python bench.py --config $c --steps $([ $c = C2 ] && echo 20 || ([ $c = C1 ] && echo 16 || echo 2)) --warmup 3 --no-cpu-baseline > gpurun_out/bench_${c}_final2.json 2> gpurun_out/bench_${c}_final2.err
done
python tools/panel_ab.py 2>&1 | tail -1
python - <<'PY'
import json
for f in ("bench_c3_final2", "bench_C1_final2", "bench_C2_final2", "bench_C5_final2"):
    try:
        j = json.load(open(f"gpurun_out/{f}.json")); r = j.get("roofline", {})
        print(f, j["value"], j["unit"], "e2e", j["e2e"]["value"], (j.get("clocks") or {}).get("sm_mhz"), "frac", r.get("frac"), j.get("refit_ms"), j.get("fp64_side"))
    except Exception as e:
        print(f, "FAILED", e)
PY
