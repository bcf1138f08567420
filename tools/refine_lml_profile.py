"""Kernel-level breakdown (torch profiler) of bo_refine and of one lock-step K7 evaluation at the C1 shape."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from bayesianoptimizer_b200 import GPEngine
n, d = int(os.environ.get("PROF_N", 3000)), 5
X = np.random.default_rng(4).random((n, d)); y = np.sin(3 * X).sum(1); y = (y - y.mean()) / y.std(ddof=1)
eng = GPEngine(torch.device("cuda", 0))
Xd, yd = torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()
eng.fit(Xd, yd, "matern52", 0.5, 1.0, 1e-3)
starts = torch.rand(10, d, dtype=torch.float64, device="cuda")
for it in (50, 50):
    torch.cuda.synchronize(); t = time.perf_counter(); eng.refine(starts, "logei", float(y.max()), iters=it); torch.cuda.synchronize()
    print("refine", it, "iters ms", (time.perf_counter() - t) * 1e3)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    eng.refine(starts, "logei", float(y.max()), iters=50); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))
R = int(os.environ.get("PROF_R", 4))
rng = np.random.default_rng(9)
th = np.concatenate([rng.uniform(np.log(0.2), np.log(2), (R, d)), np.zeros((R, 1)), rng.uniform(np.log(1e-3), np.log(1e-1), (R, 1))], axis=1)
for _ in range(3):
    torch.cuda.synchronize(); t = time.perf_counter(); eng.lml_grad_batched(Xd, yd, th); torch.cuda.synchronize()
    print("lml R =", R, "ms", (time.perf_counter() - t) * 1e3)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    eng.lml_grad_batched(Xd, yd, th); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
