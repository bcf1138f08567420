import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from bayesianoptimizer_b200 import GPEngine
eng = GPEngine(torch.device("cuda", 0))
rng = np.random.default_rng(0)
for n, N, m in ((3000, 20000, 8), (600, 20000, 8)):
    X = torch.from_numpy(rng.random((n, 5))).cuda(); Y = torch.from_numpy(rng.standard_normal((n, m))).cuda()
    Xs = torch.from_numpy(rng.random((N, 5))).cuda()
    eng.fit(X, Y[:, 0].contiguous(), "linear_matern52", 0.5, 1.0, 1e-2, linear_variance=0.3)
    for wv in (True, False):
        eng.posterior_multi(Y, Xs, with_variance=wv); torch.cuda.synchronize()
        t = time.perf_counter(); eng.posterior_multi(Y, Xs, with_variance=wv); torch.cuda.synchronize()
        print(f"n={n} N={N} m={m} with_variance={wv}: {(time.perf_counter() - t) * 1e3:.2f} ms")
    t = time.perf_counter(); eng.posterior(Xs); torch.cuda.synchronize(); print("single posterior", (time.perf_counter() - t) * 1e3, "ms")
