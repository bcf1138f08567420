#!/bin/bash
# ncu --set full of the sliced sweep kernel (pair and one-CTA variants) on a 2-wave C3-shaped pool; the raw / source
# pages are exported on the box (the .ncu-rep files are too large to bring back together)
export PROF_POOL=18944
for v in pair one; do
  if [ $v = one ]; then export BO_B200_I8_PAIR=0; fi
  python tools/profile_sweep.py > gpurun_out/plain_$v.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:sweep_i8 -s 1 -c 1 -f -o /tmp/prof_i8_$v python tools/profile_sweep.py > gpurun_out/ncu_$v.log 2>&1
  ncu -i /tmp/prof_i8_$v.ncu-rep --page raw --csv > gpurun_out/ncu_i8_${v}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_i8_$v.ncu-rep --page details > gpurun_out/ncu_i8_${v}_details.txt 2>/dev/null
  ncu -i /tmp/prof_i8_$v.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/ncu_i8_${v}_source.csv.gz
  ls -la /tmp/prof_i8_$v.ncu-rep
done
tail -n 3 gpurun_out/plain_pair.log gpurun_out/plain_one.log
du -sh gpurun_out
