"""Create / use / destroy handles repeatedly and watch the device's free memory (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesianoptimizer_b200 import GPEngine, sobol_state
torch.cuda.init(); torch.zeros(1, device="cuda")
rng = np.random.default_rng(0)
X = torch.from_numpy(rng.random((1500, 5))).cuda(); y = torch.from_numpy(rng.standard_normal(1500)).cuda()
free0 = None
for it in range(12):
    eng = GPEngine(torch.device("cuda", 0))
    eng.fit(X, y, "matern52", 0.5, 1.0, 1e-2)
    eng.sweep("ei", 0.0, sobol=sobol_state(5, 1), count=50000, topk=8)
    eng.append(X[3] * 0.5)
    eng.refine(X[:9], "logei", 0.0, iters=20)
    eng.lml_grad_batched(X, y, np.zeros((6, 7)))
    eng.topk_scores(y, 500)
    Ls = torch.eye(1500, dtype=torch.float64, device="cuda") * 0.5
    eng.load_svgp(X, y, Ls, "linear_matern52", 0.5, 1.0, 0.2, 0.0, 1e-3, 1e-4)
    eng.sweep("var", candidates=X, topk=4)
    eng.fps(X, 50, 0)
    eng.close()
    torch.cuda.synchronize()
    free, total = torch.cuda.mem_get_info()
    if it == 1: free0 = free
    print(it, "free MB", free // 2**20, flush=True)
assert free0 is not None and abs(free - free0) < 64 * 2**20, "device memory is not returned"
print("leak check ok")
