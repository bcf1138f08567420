import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesianoptimizer_b200 import GPEngine
eng = GPEngine(torch.device("cuda", 0))
for (m, n, k) in ((4096, 4096, 256), (4096, 4096, 1024), (4096, 4096, 4096), (2048, 2048, 2048)):
    row = []
    for cfg in (3, 13, 23):
        try:
            row.append(f"cfg{cfg} {eng.gemm_probe_tflops(m, n, k, cfg, 20):6.2f}")
        except Exception as e:
            row.append(f"cfg{cfg}   n/a")
    print(f"m={m} n={n} k={k}: " + "  ".join(row), flush=True)
