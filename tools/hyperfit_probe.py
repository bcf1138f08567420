"""How many lock-step evaluations does the MAP fit use, and what does a looser (reference-normalised) gtol cost in objective?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesianoptimizer_b200 import GPEngine
from bayesianoptimizer_b200 import hyperfit as hf
n, d = int(os.environ.get("PROF_N", 3000)), 5
rng = np.random.default_rng(0)
X = rng.random((n, d)); y = np.sin(3 * X).sum(1) + 0.05 * rng.standard_normal(n); y = (y - y.mean()) / y.std(ddof=1)
eng = GPEngine(torch.device("cuda", 0))
Xd, yd = torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()
lo = np.log(np.array([0.025] * d + [1e-2, 1e-4])); hi = np.log(np.array([20.0] * d + [1e2, 1.0]))
th0 = np.vstack([np.log([0.5] * d + [1.0, 1e-3]), rng.uniform(np.log(0.1), np.log(3), (3, d + 2))])
th0[1:, d + 1] = np.log(1e-2)
eng.lml_grad_batched(Xd, yd, th0)
for gtol, ftol in ((1e-5, 1e-9), (1e-5 * n, 1e-9), (1e-5 * n, 2.2e-9), (1e-4 * n, 1e-7)):
    calls = []
    def evaluate(th):
        lml, grad, status = eng.lml_grad_batched(Xd, yd, th, "matern52", 0.0)
        lml = np.asarray(lml, dtype=np.float64).copy(); grad = np.asarray(grad, dtype=np.float64).copy()
        lp, lg = hf.log_prior_and_grad(th, d, "gamma")
        calls.append(len(th))
        F = lml + lp; G = grad + lg
        bad = np.asarray(status) != 0
        F[bad] = -np.inf; G[bad] = 0
        return F, G
    torch.cuda.synchronize(); t = time.perf_counter()
    x, f, nev = hf.lbfgs_lockstep(evaluate, th0, lo, hi, maxiter=50, gtol=gtol, ftol=ftol)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t) * 1e3
    print(f"gtol={gtol:g} ftol={ftol:g}: {ms:.0f} ms, {nev} lock-step evals, restart-evals {sum(calls)}, F = {np.array2string(f, precision=4)}")
