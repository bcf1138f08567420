"""A/B of the Cholesky panel kernel (BO_B200_PANEL_FUSED=0/1): refit at n = 4096 / 8192 and one LML+gradient evaluation at n = 3000 (R = 1, 16)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesianoptimizer_b200 import GPEngine
dev = torch.device("cuda", 0)
eng = GPEngine(dev)
out = {"fused": os.environ.get("BO_B200_PANEL_FUSED", "1")}
def ev(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for n, d in ((4096, 8), (8192, 8)):
    rng = np.random.default_rng(n)
    X = torch.from_numpy(rng.random((n, d))).to(dev); y = torch.sin(3 * X).sum(1); y = (y - y.mean()) / y.std()
    out[f"fit_n{n}_ms"] = ev(lambda: eng.fit(X, y, "matern52", 0.7, 1.0, 1e-3), 10)
n, d = 3000, 5
rng = np.random.default_rng(1)
X = torch.from_numpy(rng.random((n, d))).to(dev); y = torch.sin(3 * X).sum(1); y = (y - y.mean()) / y.std()
for R in (1, 16):
    th = np.tile(np.log([0.5] * d + [1.0, 1e-2]), (R, 1)) + 0.1 * rng.standard_normal((R, d + 2))
    out[f"lml_n3000_R{R}_ms"] = ev(lambda: eng.lml_grad_batched(X, y, th, "matern52"), 10)
    l, g, s = eng.lml_grad_batched(X, y, th, "matern52")
    out[f"lml_R{R}_value0"] = float(np.asarray(l)[0])
eng.close()
print(json.dumps(out))
