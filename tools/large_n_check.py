"""One-off validation of the large-n path (n = 16384 .. 32768): fit + posterior vs the CPU oracle / self-consistency."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesianoptimizer_b200 import GPEngine, sobol_state
from oracle import gp_oracle as o
n, d = int(os.environ.get("BIG_N", 16384)), 5
X = np.random.default_rng(1).random((n, d)); y = np.sin(3 * X).sum(1) + 0.05 * np.random.default_rng(2).standard_normal(n); y = (y - y.mean()) / y.std(ddof=1)
eng = GPEngine(torch.device("cuda", 0))
Xd, yd = torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()
torch.cuda.synchronize(); t = time.perf_counter()
eng.fit(Xd, yd, "matern52", 0.5, 1.0, 1e-2)
torch.cuda.synchronize(); print("gpu fit s", time.perf_counter() - t, "mem GB", torch.cuda.mem_get_info()[0] / 1e9, flush=True)
c = np.random.default_rng(3).random((256, d)); c[:4] = X[[0, 1, n // 2, n - 1]]
t = time.perf_counter(); m, v = eng.posterior(torch.from_numpy(c).cuda()); torch.cuda.synchronize(); print("gpu posterior s", time.perf_counter() - t, flush=True)
st = sobol_state(d, 1)
t = time.perf_counter(); vals, idx = eng.sweep("logei", float(y.max()), sobol=st, count=10000, topk=4); torch.cuda.synchronize(); print("gpu sweep 1e4 s", time.perf_counter() - t, idx.tolist(), flush=True)
if os.environ.get("BIG_ORACLE", "1") == "1":
    t = time.perf_counter(); gp = o.fit(X, y, o.KERNEL_MATERN52, 0.5, 1.0, 1e-2); print("cpu fit s", time.perf_counter() - t, flush=True)
    mu, var = o.posterior(gp, c)
    em = np.abs(m.cpu().numpy() - mu) / (1e-8 * np.abs(mu) + 1e-8); ev = np.abs(v.cpu().numpy() - var) / (1e-8 * var)
    print("mean err (x tol)", em.max(), "var err (x tol)", ev.max())
