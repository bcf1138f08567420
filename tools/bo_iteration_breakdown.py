"""Wall-clock breakdown of ONE BO iteration of the drop-in class at the C1 shape (n rows of the cached CSV objective,
default GPConfig): hyper-parameter fit, final fit, pool sweep, refinement.  Development aid; bench.py is the contract."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesianoptimizer_b200.optimizer import BayesianOptimizer, GPConfig
from bayesianoptimizer_b200.simulators import DEFAULT_BOUNDS, CachedCSVSimulator

def t(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    return r, (time.perf_counter() - t0) * 1e3

n = int(os.environ.get("BO_N", 3000))
rng = np.random.default_rng(0)
lo, hi = np.array(DEFAULT_BOUNDS).T
P = rng.random((n, 5)) * (hi - lo) + lo
U = (P - lo) / (hi - lo)
Y = (np.sin(3 * U).sum(1)[:, None] + 2.5 + 0.05 * rng.standard_normal((n, 8)))
sim = CachedCSVSimulator(P, Y)
out = {}
for name, cfg in (("default", GPConfig(seed=0)), ("pool_1e4", GPConfig(seed=0, candidates_pool_size=10_000))):
    opt = BayesianOptimizer(sim, DEFAULT_BOUNDS, "/tmp/bo_iter_" + name, 0, 1, 1, gp_config=cfg)
    for i in range(n):
        opt._append_observation(P[i], Y[i], write=False)
    res = {}
    for rep in range(2):                      # second pass = warm (workspaces allocated, hyper-parameters warm-started)
        eng = opt._engine_get()
        y, _, _, _ = opt._model_targets()
        (_, res["hyperfit_ms"]) = t(lambda: opt._fit_hyperparameters(eng, opt.train_X, y))
        gp, res["fit_gp_model_ms"] = t(opt.fit_gp_model)
        x, res["suggest_ms"] = t(lambda: opt.suggest(1, gp))
        (v, i), res["sweep_ms"] = t(lambda: eng.sweep(cfg.acquisition, opt._best_f(), sobol=__import__("bayesianoptimizer_b200").sobol_state(5, 1),
                                                      count=cfg.candidates_pool_size, topk=10))
        st = eng.sobol_points(__import__("bayesianoptimizer_b200").sobol_state(5, 1), i)
        _, res["refine_ms"] = t(lambda: eng.refine(st, cfg.acquisition, opt._best_f(), iters=cfg.refine_iters))
        _, res["plain_fit_ms"] = t(lambda: eng.fit(opt.train_X, y, cfg.kernel, *opt._hyper[:3]))
    out[name] = {k: round(v, 3) for k, v in res.items()}
    opt.close()
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bo_iteration.json", "w"), indent=1)
