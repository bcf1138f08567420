"""Small fixed workload for ncu: fit at C3 shape + one 2-wave sweep (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesianoptimizer_b200 import GPEngine, sobol_state
n, d = int(os.environ.get("PROF_N", 4096)), int(os.environ.get("PROF_D", 8))
N = int(os.environ.get("PROF_POOL", 148 * 128 * 2))
X = np.random.default_rng(4).random((n, d))
y = np.sin(3.0 * X).sum(axis=1) + 0.05 * np.random.default_rng(5).standard_normal(n)
y = (y - y.mean()) / y.std(ddof=1)
eng = GPEngine(torch.device("cuda", 0))
eng.fit(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(), "matern52", 0.7, 1.0, 1e-3)
st = sobol_state(d, 6)
for _ in range(2):
    v, i = eng.sweep("ei", float(y.max()), sobol=st, count=N, topk=1)
torch.cuda.synchronize()
print("sweep ms", eng.last_sweep_ms(), "argmax", i.item(), v.item())
