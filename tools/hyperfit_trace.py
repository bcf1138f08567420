"""Trace of a WARM-started MAP fit: optimum on n-10 points, then 10 more points and a refit from that optimum."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesianoptimizer_b200 import GPEngine
from bayesianoptimizer_b200 import hyperfit as hf
n, d = int(os.environ.get("PROF_N", 3000)), 5
rng = np.random.default_rng(0)
X = rng.random((n, d)); y = np.sin(3 * X).sum(1) + 0.05 * rng.standard_normal(n); y = (y - y.mean()) / y.std(ddof=1)
eng = GPEngine(torch.device("cuda", 0))
lo = np.log(np.array([0.025] * d + [1e-2, 1e-4])); hi = np.log(np.array([20.0] * d + [1e2, 1.0]))
def run(Xd, yd, th0, label, **kw):
    trace = []
    def evaluate(th):
        lml, grad, status = eng.lml_grad_batched(Xd, yd, th, "matern52", 0.0)
        lml = np.asarray(lml, dtype=np.float64).copy(); grad = np.asarray(grad, dtype=np.float64).copy()
        lp, lg = hf.log_prior_and_grad(th, d, "gamma")
        F = lml + lp; G = grad + lg
        trace.append((len(th), float(F.max()), float(np.abs(G).max())))
        return F, G
    torch.cuda.synchronize(); t = time.perf_counter()
    x, f, nev = hf.lbfgs_lockstep(evaluate, th0, lo, hi, maxiter=50, **kw)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t) * 1e3
    print(f"{label}: {ms:.0f} ms, {nev} evals, F={f.max():.6f}")
    print("   trace (R, maxF, max|G|):", [(r, round(F, 4), float(f"{g:.2g}")) for r, F, g in trace])
    return x[int(np.argmax(f))]
m = n - 10
Xa, ya = torch.from_numpy(X[:m]).cuda(), torch.from_numpy(y[:m]).cuda()
Xb, yb = torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()
th_cold = np.log(np.array([[0.5] * d + [1.0, 1e-3]]))
opt_a = run(Xa, ya, th_cold, "cold, n-10")
run(Xb, yb, opt_a[None, :], "warm, n (1 restart)")
run(Xb, yb, opt_a[None, :], "warm, n, fscale=n", fscale=float(n))
run(Xa, ya, th_cold, "cold, n-10, fscale=n", fscale=float(n))
th4 = np.vstack([opt_a, rng.uniform(np.log(0.1), np.log(3), (3, d + 2))]); th4[1:, d + 1] = np.log(1e-2)
run(Xb, yb, th4, "warm + 3 random, fscale=1")
run(Xb, yb, th4, "warm + 3 random, fscale=n", fscale=float(n))
