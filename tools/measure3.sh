#!/bin/bash
for n in 2048 3072 4096; do for pair in 1 0; do echo "n=$n pair=$pair S=7w"; BO_B200_I8_PAIR=$pair timeout 200 python tools/i8_sweep_check.py $n 8 1200000 7 2>&1 | grep -E "i8: sweep"; done; done
BO_B200_SWEEP_FLAGS=6 timeout 300 python tools/i8_sweep_check.py 4096 8 2400000 7 2>&1 | grep -E "MMA issuer|producer|i8: sweep"
export PROF_POOL=18944
python tools/profile_sweep.py > gpurun_out/plain_pair.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sweep_i8 -s 1 -c 1 -f -o /tmp/prof_i8_pair python tools/profile_sweep.py > gpurun_out/ncu_pair.log 2>&1
ncu -i /tmp/prof_i8_pair.ncu-rep --page raw --csv > gpurun_out/ncu_i8_pair7_raw.csv 2>/dev/null
ncu -i /tmp/prof_i8_pair.ncu-rep --page details > gpurun_out/ncu_i8_pair7_details.txt 2>/dev/null
tail -1 gpurun_out/plain_pair.log; grep -c . gpurun_out/ncu_i8_pair7_raw.csv
python bench.py --config C2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_C2_w8.json 2>/dev/null; cut -c1-150 gpurun_out/bench_C2_w8.json
