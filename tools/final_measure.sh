#!/bin/bash
# round-2 final measurements: bench lines (all configs), ncu launch list of the default bench command, full ncu capture of
# the dominant kernel (2-wave pool: 148 x 64 x 2 candidates at the C3 shape)
set -x
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c3_final.json 2> gpurun_out/bench_c3_final.err
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_c3_reference.json 2> gpurun_out/bench_c3_reference.err
python bench.py --config C2 --steps 20 --warmup 5 > gpurun_out/bench_C2_final.json 2> gpurun_out/bench_C2_final.err
python bench.py --config C5 --steps 3 --warmup 3 > gpurun_out/bench_C5_final.json 2> gpurun_out/bench_C5_final.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
export PROF_POOL=18944
python tools/profile_sweep.py > gpurun_out/plain_pair.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sweep_i8 -s 1 -c 1 -f -o /tmp/prof_i8_pair python tools/profile_sweep.py > gpurun_out/ncu_pair.log 2>&1
ncu -i /tmp/prof_i8_pair.ncu-rep --page raw --csv > gpurun_out/ncu_i8_pair_raw.csv 2>/dev/null
ncu -i /tmp/prof_i8_pair.ncu-rep --page details > gpurun_out/ncu_i8_pair_details.txt 2>/dev/null
tail -n 2 gpurun_out/plain_pair.log
cut -c1-600 gpurun_out/bench_c3_final.json; echo; cut -c1-400 gpurun_out/bench_c3_reference.json; echo; cut -c1-400 gpurun_out/bench_C2_final.json; echo; cut -c1-300 gpurun_out/bench_C5_final.json
