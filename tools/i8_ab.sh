#!/bin/bash
# A/B of two builds of libbo_b200.so on the sliced sweep (run on the GPU box from the repo root):
#   build the version to compare against out of tree (copy bayesianoptimizer_b200/csrc + include/ to a scratch directory at the
#   older commit, `make` there) and drop its library next to the current one as bayesianoptimizer_b200/libbo_b200_prev.so
#   (git-ignored, travels with gpurun); this script times the current build, swaps the file in the box's scratch copy and
#   times the other one.  profiles/r01_i8_builder_overlap_ab.log is the output for "panel builders on their own warpgroup"
#   against "panel build serial with the contraction".
set -e
[ -f bayesianoptimizer_b200/libbo_b200_prev.so ] || { echo "bayesianoptimizer_b200/libbo_b200_prev.so missing (see the header of this script)"; exit 1; }
run() { for cfg in "512 5 400000" "1024 5 400000" "2048 8 200000" "4096 8 400000"; do timeout 200 python tools/i8_sweep_check.py $cfg 7 2>&1 | grep -E "^i8:|^fp64:" | tr '\n' ' '; echo " [$cfg]"; done; }
echo "== new"; run
cp bayesianoptimizer_b200/libbo_b200_prev.so bayesianoptimizer_b200/libbo_b200.so
echo "== prev"; run
