#!/bin/bash
for lib in "$@"; do
  echo "=== $lib (pair)"
  BO_B200_LIB=$PWD/bayesianoptimizer_b200/$lib BO_B200_SWEEP_FLAGS=6 timeout 300 python tools/i8_sweep_check.py 4096 8 2400000 8 2>&1 | grep -E "sweep_i8|i8: sweep|var   max|top-k|timed out|rror"
done
