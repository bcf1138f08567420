"""Sliced sweep -> bo_release_workspace -> sliced sweep: the int8 operand pack and panels are re-created on demand and the
result is bit-identical.  python tools/i8_release_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesianoptimizer_b200 import GPEngine, sobol_state

X = np.random.default_rng(1).random((700, 6)); y = np.sin(3 * X).sum(1); y = (y - y.mean()) / y.std(ddof=1)
eng = GPEngine(torch.device("cuda", 0))
eng.fit(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(), "matern52", 0.6, 1.0, 1e-3)
st = sobol_state(6, 5)
v1, i1 = eng.sweep("ei", float(y.max()), sobol=st, count=40_000, topk=8)
assert eng.last_sweep_path() == 8
eng.release_workspace()
v2, i2 = eng.sweep("ei", float(y.max()), sobol=st, count=40_000, topk=8)
assert eng.last_sweep_path() == 8 and torch.equal(v1, v2) and torch.equal(i1, i2)
eng.close()
print("release check ok", i1.tolist())
