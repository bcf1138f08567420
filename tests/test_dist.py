"""CPU tier: candidate sharding + the single (value, index) exchange, incl. a world_size-2 gloo run."""
import os
import subprocess
import sys

import numpy as np
import torch

from bayesianoptimizer_b200.dist import merge_topk, shard_range
from oracle import gp_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition_the_pool():
    for total in (0, 1, 7, 128, 10_000_000, 10_000_001):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert sum(c for _, c in spans) == total
            pos = 0
            for first, count in spans:
                if count:
                    assert first == pos
                pos += count


def test_merge_matches_oracle_order_with_ties_and_empty_slots():
    rng = np.random.default_rng(0)
    for G in (1, 2, 5, 8):
        k = 6
        vals = rng.integers(0, 4, size=(G, k)).astype(np.float64)      # many exact ties
        idx = rng.permutation(G * k).reshape(G, k).astype(np.int64)
        vals[-1, -2:] = -np.inf; idx[-1, -2:] = -1                      # empty slots of a short shard
        mv, mi = merge_topk(torch.from_numpy(vals), torch.from_numpy(idx), k)
        ov, oi = o.merge_topk(list(vals), list(idx), k)
        m = len(oi)
        assert mi.tolist()[:m] == oi.tolist() and mv.tolist()[:m] == ov.tolist()
        assert mi.tolist()[m:] == [-1] * (k - m)
    mv, mi = merge_topk(torch.tensor([[-np.inf]]), torch.tensor([[-1]]), 3)
    assert mi.tolist() == [-1, -1, -1]


_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from oracle_engine import OracleEngine
from bayesianoptimizer_b200 import sobol_state
from bayesianoptimizer_b200.dist import sharded_sweep
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
X = np.random.default_rng(1).random((60, 3)); y = np.sin(3 * X).sum(1)
eng = OracleEngine().fit(X, y, "matern52", 0.5, 1.0, 1e-3)
sob = sobol_state(3, 7)
v, i = sharded_sweep(eng, "ucb", 0.0, 2.0, sob, 4001, 5, rank, world)
v1, i1 = eng.sweep("ucb", 0.0, 2.0, sobol=sob, count=4001, topk=5)
assert i.tolist() == i1.tolist() and v.tolist() == v1.tolist(), (rank, i.tolist(), i1.tolist())
# restart sharding of the batched LML (BASELINE config 5): 5 restarts over 2 ranks == one unsharded call
from bayesianoptimizer_b200.dist import sharded_lml_grad
th = np.random.default_rng(3).uniform(-1.0, 0.5, (5, 5)); th[:, 4] = np.log(1e-2)
l, g, st = sharded_lml_grad(eng, X, y, th, "matern52", 0.0, rank, world)
l1, g1, st1 = eng.lml_grad_batched(X, y, th, "matern52", 0.0)
assert l.shape == (5,) and g.shape == (5, 5) and st.tolist() == st1.tolist()
assert torch.equal(l, l1.to(torch.float64)) and torch.equal(g, g1.to(torch.float64)), (rank, l, l1)
# task sharding of a batched SVGP scan: partial variance sums meet in one all-reduce
from bayesianoptimizer_b200.dist import task_shard, allreduce_score
from oracle import gp_oracle as o
rng = np.random.default_rng(5)
tasks = []
for t in range(3):
    M = 20
    Ls = np.tril(rng.standard_normal((M, M)) * 0.05) + np.diag(0.3 + 0.5 * rng.random(M))
    tasks.append(o.SVGPTask(rng.standard_normal((M, 3)), o.KERNEL_LINEAR_MATERN52, rng.uniform(0.5, 1.5, 3), 1.2, 0.2, 0.0, 1e-3, 1e-6,
                            rng.standard_normal(M), Ls))
xs = rng.standard_normal((200, 3))
mine = task_shard(3, rank, world)
part = torch.zeros(200, dtype=torch.float64)
for t in mine:
    e = OracleEngine().load_svgp(tasks[t].Z, tasks[t].m, tasks[t].Ls, "linear_matern52", tasks[t].lengthscale, 1.2, 0.2, 0.0, 1e-3, 1e-6)
    part += e.sweep("var", candidates=xs, topk=0, return_all=True)[3]
total = allreduce_score(part)
np.testing.assert_allclose(total.numpy(), o.svgp_variance_score(tasks, xs), rtol=1e-12)
assert sorted(sum((task_shard(3, r, world) for r in range(world)), [])) == [0, 1, 2]
print("rank", rank, "ok", i.tolist())
dist.destroy_process_group()
'''


def test_world_size_2_gloo_sharded_sweep_equals_single(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", str(script), ROOT],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2


_OPT_WORKER = r'''
import os, sys, json, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from oracle_engine import OracleEngine
from bayesianoptimizer_b200.optimizer import BayesianOptimizer, GPConfig
from bayesianoptimizer_b200.simulators import DEFAULT_BOUNDS, CachedCSVSimulator
dist.init_process_group("gloo")
rank = dist.get_rank()
z = np.load(os.path.join(sys.argv[1], "tests", "golden", "csv_cache_rows.npz"))
sim = CachedCSVSimulator(z["params"], z["outputs"])
cfg = GPConfig(candidates_pool_size=600, num_restarts=3, refine_iters=4, hyper_restarts=4, hyper_refine=2, hyper_maxiter=3, seed=None)
opt = BayesianOptimizer(sim, DEFAULT_BOUNDS, sys.argv[2], n_initial_points=12, n_batches=2, batch_size=2, gp_config=cfg,
                        engine_factory=OracleEngine, device=torch.device("cpu"))
opt.optimize()
out = {"X": opt.train_X.numpy().tolist(), "Y": opt.train_Y.numpy().tolist(), "hyper": [np.asarray(h).tolist() for h in opt._hyper],
       "sim_calls": sim.calls}
gathered = [None, None]
dist.all_gather_object(gathered, out)
assert gathered[0]["X"] == gathered[1]["X"] and gathered[0]["Y"] == gathered[1]["Y"], "replicas diverged (data)"
assert gathered[0]["hyper"] == gathered[1]["hyper"], "replicas diverged (hyper-parameters)"
assert gathered[0]["sim_calls"] == 16 and gathered[1]["sim_calls"] == 0, "the simulator must run on rank 0 only"
dist.barrier()
if rank == 0:
    rows = open(opt.results_file).read().strip().split("\n")
    assert len(rows) == 1 + 16, len(rows)          # header + one row per evaluation, written once
print("rank", rank, "ok")
dist.destroy_process_group()
'''


def test_world_size_2_driver_mode_is_replica_consistent_with_seed_none(tmp_path):
    """INTEGRATION.md multi-GPU driver mode: two ranks run BayesianOptimizer.optimize with GPConfig.seed = None; LHS points,
    hyper-parameter restarts and suggestions must agree on both, the simulator and the CSV belong to rank 0."""
    script = tmp_path / "opt_worker.py"
    script.write_text(_OPT_WORKER)
    out = tmp_path / "run"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29741", str(script), ROOT, str(out)],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2
