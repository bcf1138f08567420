"""BASELINE config 5 (256 restarts, n=2048, d=10) with the restarts sharded over the ranks (torchrun, NCCL)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from bayesianoptimizer_b200 import GPEngine
from bayesianoptimizer_b200.dist import sharded_lml_grad
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = GPEngine(torch.device("cuda", local))
n, d, R = 2048, 10, 256
X = np.random.default_rng(8).random((n, d)); y = np.sin(3 * X).sum(1) + 0.05 * np.random.default_rng(5).standard_normal(n); y = (y - y.mean()) / y.std(ddof=1)
Xd, yd = torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()
rng = np.random.default_rng(9)
th = np.concatenate([rng.uniform(np.log(0.05), np.log(5), (R, d)), np.zeros((R, 1)), rng.uniform(np.log(1e-4), np.log(1e-1), (R, 1))], axis=1)
out = sharded_lml_grad(eng, Xd, yd, th, "matern52", 0.0, rank, world)
ts = []
for _ in range(3):
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t = time.perf_counter()
    lml, grad, st = sharded_lml_grad(eng, Xd, yd, th, "matern52", 0.0, rank, world)
    torch.cuda.synchronize(); ts.append((time.perf_counter() - t) * 1e3)
if world > 1:
    tt = torch.tensor([min(ts)], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms = tt.item()
else:
    ms = min(ts)
if rank == 0:
    print(json.dumps({"config": "C5", "n": n, "d": d, "R": R, "n_gpus": world, "ms_per_evaluation": ms, "restarts_per_s": R / ms * 1e3,
                      "tflops": R * float(n) ** 3 / ms * 1e-9, "checksum": float(lml.sum()), "failed": int((st != 0).sum())}))
if world > 1:
    dist.destroy_process_group()
