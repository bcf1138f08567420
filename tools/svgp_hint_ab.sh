#!/bin/bash
for F in 2 74 2 74; do
  echo "flags=$F"; BO_B200_SWEEP_FLAGS=$F timeout 300 python tools/svgp_scan_timing.py 2>&1 | grep -E "auto_scan_ms|auto_scan_topk"
done
