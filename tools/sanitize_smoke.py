"""Small end-to-end pass over every entry point for compute-sanitizer (kept tiny: sanitizer is ~50x slower)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesianoptimizer_b200 import GPEngine, sobol_state
rng = np.random.default_rng(0)
eng = GPEngine(torch.device("cuda", 0))
for (n, d, kern) in ((200, 5, "matern52"), (129, 3, "rbf")):
    X = rng.random((n, d)); y = np.sin(3 * X).sum(1); y = (y - y.mean()) / y.std(ddof=1)
    Xd, yd = torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()
    eng.fit(Xd, yd, kern, 0.5, 1.0, 1e-3)
    st = sobol_state(d, 1)
    v, i, m, s, a = eng.sweep("logei", float(y.max()), sobol=st, count=300, topk=5, return_all=True)
    c = torch.rand(140, d, dtype=torch.float64, device="cuda")
    eng.posterior(c)
    eng.sweep_host("ei", 0.1, candidates=c.cpu().numpy(), topk=3)
    x = eng.sobol_points(st, i)
    eng.acq_grad(x, "ei", float(y.max()))
    eng.refine(x, "logei", float(y.max()), iters=3)
    eng.append(x[0]); eng.append(x[1], 0.3)
    eng.sweep("ucb", 0.0, 2.0, sobol=st, count=200, topk=2)
    th = np.log(np.concatenate([np.full((3, d), 0.5), np.ones((3, 1)), np.full((3, 1), 1e-2)], axis=1))
    eng.lml_grad_batched(Xd, yd, th, kern)
    eng.state()
torch.cuda.synchronize()
eng.close()
print("sanitize smoke done")
