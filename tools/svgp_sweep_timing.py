"""Timing of the SVGP pool sweep at Bayesian7's shape (T tasks x M = 2048 inducing points, 10^4-candidate pool, linear + Matern-5/2):
FP64 two-pass kernel vs the sliced one-pass kernel over [L^-1; Ls^T L^-1].  Development aid; writes gpurun_out/svgp_sweep_timing.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesianoptimizer_b200 import GPEngine

M, d = int(os.environ.get("SV_M", 2048)), 5
rng = np.random.default_rng(0)
Z = rng.standard_normal((M, d))
Ls = np.tril(rng.standard_normal((M, M)) * 0.05 / np.sqrt(M / 64)) + np.diag(0.3 + 0.5 * rng.random(M))
m = rng.standard_normal(M)
dev = torch.device("cuda", 0)
eng = GPEngine(dev)
c = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
t0 = time.perf_counter()
eng.load_svgp(c(Z), c(m), c(Ls), "linear_matern52", rng.uniform(0.8, 2.0, d), 1.3, 0.2, 0.1, 2e-3, 1e-4)
torch.cuda.synchronize()
out = {"M": M, "load_ms_first": (time.perf_counter() - t0) * 1e3}
t0 = time.perf_counter()
eng.load_svgp(c(Z), c(m), c(Ls), "linear_matern52", rng.uniform(0.8, 2.0, d), 1.3, 0.2, 0.1, 2e-3, 1e-4)
torch.cuda.synchronize()
out["load_ms_warm"] = (time.perf_counter() - t0) * 1e3
for N in (10_000, 100_000, 1_000_000):
    xs = c(rng.standard_normal((N, d)))
    for mode in ("fp64", "i8x7", "i8x8", "auto"):
        eng.set_sweep_mode(mode)
        for _ in range(2):
            eng.sweep("var", candidates=xs, topk=8, min_variance=1e-3)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 5 if N <= 100_000 else 2
        ks = []
        for _ in range(reps):
            eng.sweep("var", candidates=xs, topk=8, min_variance=1e-3)
            ks.append(eng.last_sweep_ms())
        torch.cuda.synchronize()
        out[f"N{N}_{mode}"] = {"wall_ms": (time.perf_counter() - t0) * 1e3 / reps, "kernel_ms": float(np.mean(ks)), "path": eng.last_sweep_path(),
                              "flagged": eng.last_sweep_flagged()}
eng.close()
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/svgp_sweep_timing.json", "w"), indent=1)
