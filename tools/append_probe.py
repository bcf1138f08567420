"""bo_append at n = 4096 / 8176: per-call time (CUDA events) -- run under ncu for the per-kernel list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesianoptimizer_b200 import GPEngine
dev = torch.device("cuda", 0)
eng = GPEngine(dev)
for n in (4096, 8176):
    d = 8
    rng = np.random.default_rng(n)
    X = torch.from_numpy(rng.random((n, d))).to(dev); y = torch.sin(3 * X).sum(1); y = (y - y.mean()) / y.std()
    eng.fit(X, y, "matern52", 0.7, 1.0, 1e-3)
    pts = torch.rand(24, d, dtype=torch.float64, device=dev)
    for j in range(4): eng.append(pts[j])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for j in range(16): eng.append(pts[4 + j])
    e1.record(); torch.cuda.synchronize()
    print(n, "append ms", e0.elapsed_time(e1) / 16)
eng.close()
