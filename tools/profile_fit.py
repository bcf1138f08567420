import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesianoptimizer_b200 import GPEngine
n, d = int(os.environ.get("PROF_N", 4096)), 8
X = np.random.default_rng(4).random((n, d)); y = np.sin(3 * X).sum(1); y = (y - y.mean()) / y.std(ddof=1)
eng = GPEngine(torch.device("cuda", 0))
eng.fit(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(), "matern52", 0.7, 1.0, 1e-3)
torch.cuda.synchronize(); print("ok")
