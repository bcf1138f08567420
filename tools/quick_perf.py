"""Quick device timings of fit and sweep (development aid; bench.py is the contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesianoptimizer_b200 import GPEngine, sobol_state

def synth(n, d, sx, sy):
    X = np.random.default_rng(sx).random((n, d))
    y = np.sin(3.0 * X).sum(axis=1) + 0.05 * np.random.default_rng(sy).standard_normal(n)
    return X, (y - y.mean()) / y.std(ddof=1)

eng = GPEngine(torch.device("cuda", 0))
print("fp64 peak dmma", eng.fp64_peak_tflops(True, 0.3), "dfma", eng.fp64_peak_tflops(False, 0.3))
for (n, d, ls, N) in ((512, 5, 0.5, 1_000_000), (4096, 8, 0.7, 148 * 128 * 8), (8192, 8, 0.7, 148 * 128 * 2)):
    X, y = synth(n, d, 4, 5)
    Xd, yd = torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()
    for _ in range(2):
        torch.cuda.synchronize(); t = time.perf_counter()
        eng.fit(Xd, yd, "matern52", ls, 1.0, 1e-3)
        torch.cuda.synchronize(); fit_ms = (time.perf_counter() - t) * 1e3
    st = sobol_state(d, 6)
    for rep in range(2):
        v, i = eng.sweep("ei", float(y.max()), sobol=st, count=N, topk=1)
        torch.cuda.synchronize()
        ms = eng.last_sweep_ms()
    flop = N * (n * n + n * (3 * d + 12.0))
    print(f"n={n} d={d} fit {fit_ms:.2f} ms | sweep N={N}: {ms:.2f} ms -> {N / ms * 1e3:.4g} cand/s, {flop / ms * 1e-9:.2f} TFLOP/s (alg)")
