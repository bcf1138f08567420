import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesianoptimizer_b200 import GPEngine
eng = GPEngine(torch.device("cuda", 0))
m, n, k, cfg = (int(v) for v in sys.argv[1:5])
print(m, n, k, cfg, eng.gemm_probe_tflops(m, n, k, cfg, 3))
