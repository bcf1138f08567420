import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesianoptimizer_b200 import GPEngine
eng = GPEngine(torch.device("cuda", 0))
for n, d in ((3000, 5), (2048, 10), (1000, 5)):
    X = np.random.default_rng(8).random((n, d)); y = np.sin(3 * X).sum(1); y = (y - y.mean()) / y.std(ddof=1)
    Xd, yd = torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()
    rng = np.random.default_rng(9)
    for R in (1, 2, 4, 16, 64):
        th = np.concatenate([rng.uniform(np.log(0.2), np.log(2), (R, d)), np.zeros((R, 1)), rng.uniform(np.log(1e-3), np.log(1e-1), (R, 1))], axis=1)
        row = []
        for G in ("1", "2", "4"):
            os.environ["BO_B200_LML_GROUPS"] = G
            eng.lml_grad_batched(Xd, yd, th)
            ts = []
            for _ in range(3):
                torch.cuda.synchronize(); t = time.perf_counter(); out = eng.lml_grad_batched(Xd, yd, th); torch.cuda.synchronize()
                ts.append((time.perf_counter() - t) * 1e3)
            row.append(f"G={G}: {min(ts):7.2f} ms")
        print(f"n={n} d={d} R={R}: " + "  ".join(row), flush=True)
