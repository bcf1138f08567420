#!/bin/bash
# A/B of sweep builds on the C3 shape: sustained rate + in-kernel wait accounting (CTA 0) for each library variant
for lib in "$@"; do
  echo "=== $lib"
  BO_B200_LIB=$PWD/bayesianoptimizer_b200/$lib BO_B200_SWEEP_FLAGS=6 python tools/i8_sweep_check.py 4096 8 2400000 8 2>&1 | grep -E "sweep_i8 CTA|i8: sweep|var   max|top-k"
done
