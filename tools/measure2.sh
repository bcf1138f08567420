#!/bin/bash
python -m pytest tests/test_gpu_i8.py -q -m gpu -x 2>&1 | tail -4
python bench.py --config C2 --steps 20 --warmup 5 > gpurun_out/bench_C2_final.json 2> gpurun_out/bench_C2_final.err; cut -c1-300 gpurun_out/bench_C2_final.json; echo
for pair in 1 0; do echo "n=2048 pair=$pair"; BO_B200_I8_PAIR=$pair timeout 200 python tools/i8_sweep_check.py 2048 8 1200000 8 2>&1 | grep -E "i8: sweep"; done
for pair in 1 0; do echo "n=3072 pair=$pair"; BO_B200_I8_PAIR=$pair timeout 200 python tools/i8_sweep_check.py 3072 8 1200000 8 2>&1 | grep -E "i8: sweep"; done
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sweep|topk|i8_" -c 200 --csv --log-file gpurun_out/launches_bench_sweeps.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
