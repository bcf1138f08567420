#!/bin/bash
export BO_B200_LIB=$PWD/bayesianoptimizer_b200/libbo_b200_dbg.so
echo "--- n=1000 d=5 N=200000 S=8 (pair, variant build)"; timeout 120 python tools/i8_sweep_check.py 1000 5 200000 8 2>&1 | tail -8
echo "--- C3 shape, pair, accounting"; BO_B200_SWEEP_FLAGS=6 timeout 300 python tools/i8_sweep_check.py 4096 8 2400000 8 2>&1 | grep -E "sweep_i8|i8: sweep|var   max|top-k|timed out|rror"
