#!/bin/bash
export BO_I8_WAIT_CYCLES=4000000000
timeout 900 python -m pytest tests/test_gpu_svgp.py -q 2>&1 | tail -25
timeout 300 python tools/svgp_scan_timing.py 2>&1 | tail -12
