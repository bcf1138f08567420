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
