"""Per-kernel time of one fit via torch profiler (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, time
from bayesianoptimizer_b200 import GPEngine
n, d = int(os.environ.get("PROF_N", 4096)), 8
X = np.random.default_rng(4).random((n, d)); y = np.sin(3 * X).sum(1); y = (y - y.mean()) / y.std(ddof=1)
eng = GPEngine(torch.device("cuda", 0))
Xd, yd = torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()
for _ in range(3):
    torch.cuda.synchronize(); t = time.perf_counter(); eng.fit(Xd, yd, "matern52", 0.7, 1.0, 1e-3); torch.cuda.synchronize()
    print("fit ms", (time.perf_counter() - t) * 1e3)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    eng.fit(Xd, yd, "matern52", 0.7, 1.0, 1e-3); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
