#!/bin/bash
BO_B200_PANEL_FUSED=0 python tools/panel_ab.py 2>&1 | tail -1
BO_B200_PANEL_FUSED=1 python tools/panel_ab.py 2>&1 | tail -1
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_extras.py tests/test_gpu_n4.py tests/test_gpu_fuzz.py tests/test_gpu_optimizer.py -q -x 2>&1 | tail -6
