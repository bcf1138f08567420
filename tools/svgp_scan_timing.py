"""Bayesian7's pool scan through BatchSVGPPredictor (T = 8 tasks, M = 2048, 10^4-candidate pool, top-8000, FPS-500) per sweep mode."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesianoptimizer_b200.simulators import DEFAULT_BOUNDS
from bayesianoptimizer_b200.svgp import BatchSVGPPredictor, SVGPTaskState

T, M, d, N = 8, 2048, 5, 10_000
dev = torch.device("cuda", 0)
c = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
rng = np.random.default_rng(0)
states = []
for k in range(T):
    Z = rng.standard_normal((M, d))
    Ls = np.tril(rng.standard_normal((M, M)) * 0.05 / np.sqrt(M / 64)) + np.diag(0.3 + 0.5 * rng.random(M))
    states.append(SVGPTaskState(c(Z), c(rng.standard_normal(M)), c(Ls), torch.from_numpy(rng.uniform(0.8, 2.0, d)), 1.3, 0.2, 0.1, 2e-3))
bounds = torch.tensor(DEFAULT_BOUNDS, dtype=torch.float64).t()
pred = BatchSVGPPredictor(dev, states, jitter=1e-4, bounds=bounds, x_log_mean=rng.standard_normal((1, d)), x_log_std=rng.uniform(0.5, 2.0, (1, d)))
U = c(rng.random((N, d)))
out = {}
for mode in ("fp64", "auto"):
    for e in pred.engines:
        e.set_sweep_mode(mode)
    for _ in range(2):
        pred.variance_score(U, min_variance=1e-3); pred.select_batch(U, 500)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        pred.variance_score(U, min_variance=1e-3)
    torch.cuda.synchronize(); out[mode + "_scan_ms"] = (time.perf_counter() - t0) * 1e3 / 5
    t0 = time.perf_counter()
    for _ in range(5):
        pred.select_batch(U, 500)
    torch.cuda.synchronize(); out[mode + "_scan_topk_fps_ms"] = (time.perf_counter() - t0) * 1e3 / 5
    out[mode + "_path"] = pred.engines[0].last_sweep_path()
pred.close()
print(json.dumps(out, indent=1))
json.dump(out, open("gpurun_out/svgp_scan_timing.json", "w"), indent=1)
