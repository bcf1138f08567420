"""Small fixed workload for ncu: one SVGP task (M inducing points) loaded + one sweep over a few waves (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesianoptimizer_b200 import GPEngine
M, d = int(os.environ.get("PROF_M", 2048)), int(os.environ.get("PROF_D", 5))
N = int(os.environ.get("PROF_POOL", 148 * 128 * 4))
g = np.random.default_rng(21)
Ls = np.tril(g.standard_normal((M, M)) * 0.01) + np.diag(0.3 + 0.5 * g.random(M))
eng = GPEngine(torch.device("cuda", 0))
eng.load_svgp(torch.from_numpy(g.standard_normal((M, d))).cuda(), torch.from_numpy(g.standard_normal(M)).cuda(), torch.from_numpy(Ls).cuda(),
              "linear_matern52", 1.5, 1.0, 0.2, 0.0, 1e-3, 1e-4)
U = torch.randn(N, d, dtype=torch.float64, device="cuda")
for _ in range(2):
    eng.sweep("var", candidates=U, topk=8, min_variance=1e-3)
torch.cuda.synchronize()
print("svgp sweep ms", eng.last_sweep_ms(), "TFLOP/s", N * (2.0 * M * M + M * (3 * d + 14)) / eng.last_sweep_ms() * 1e-9)
