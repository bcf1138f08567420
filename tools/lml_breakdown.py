import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, time
from bayesianoptimizer_b200 import GPEngine
n, d, R = 2048, 10, 32
X = np.random.default_rng(8).random((n, d)); y = np.sin(3 * X).sum(1); y = (y - y.mean()) / y.std(ddof=1)
eng = GPEngine(torch.device("cuda", 0))
Xd, yd = torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()
rng = np.random.default_rng(9)
th = np.concatenate([rng.uniform(np.log(0.05), np.log(5), (R, d)), np.zeros((R, 1)), rng.uniform(np.log(1e-4), np.log(1e-1), (R, 1))], axis=1)
eng.lml_grad_batched(Xd, yd, th)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    eng.lml_grad_batched(Xd, yd, th); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=50))
