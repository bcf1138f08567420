#!/bin/bash
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
timeout 300 python tools/svgp_scan_timing.py 2>&1 | grep -E "auto_scan|fp64_scan"
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c3_final.json 2> gpurun_out/bench_c3_final.err
python -c "
import json; j=json.load(open('gpurun_out/bench_c3_final.json')); r=j['roofline']; print('C3', j['value'], 'e2e', j['e2e']['value'], j['clocks'], r['frac'], r.get('frac_of_sustained_peak'), r.get('traffic'), j.get('cpu_baseline',{}).get('value'), j.get('argmax_check_fp64_full_pool'))"
