run() { for cfg in "512 5 400000" "1024 5 400000" "2048 8 200000" "4096 8 400000"; do timeout 200 python tools/i8_sweep_check.py $cfg 7 2>&1 | grep -E "^i8:|^fp64:" | tr '\n' ' '; echo " [$cfg]"; done; }
echo "== new (builder warps)"; run
cp bayesianoptimizer_b200/libbo_b200_prev.so bayesianoptimizer_b200/libbo_b200.so
echo "== prev (serial phase A)"; run
