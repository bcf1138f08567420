"""Timings the bench lines do not carry (missing #5), drop-in class at the C1 shapes:
cold multi-start MAP fit (16 screened restarts, 4 refined) and warm single-start refit at n = 3000 and n = 4096, and one
q = 1000 suggestion (top-K_big -> device FPS, the shape of Bayesian7.py:676-688 with main.py:15's batch_size) next to a
q = 16 Kriging-believer batch with the reference's 10^4-candidate pool.   python tools/driver_timings.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesianoptimizer_b200.optimizer import BayesianOptimizer, GPConfig
from bayesianoptimizer_b200.simulators import DEFAULT_BOUNDS, CachedCSVSimulator

def t(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    return r, (time.perf_counter() - t0) * 1e3

out = {}
for n in (3000, 4096):
    rng = np.random.default_rng(0)
    lo, hi = np.array(DEFAULT_BOUNDS).T
    P = rng.random((n, 5)) * (hi - lo) + lo
    U = (P - lo) / (hi - lo)
    Y = (np.sin(3 * U).sum(1)[:, None] + 2.5 + 0.05 * rng.standard_normal((n, 8)))
    sim = CachedCSVSimulator(P, Y)
    res = {}
    opt = BayesianOptimizer(sim, DEFAULT_BOUNDS, f"/tmp/bo_drv_{n}", 0, 1, 1, gp_config=GPConfig(seed=0, candidates_pool_size=10_000))
    for i in range(n):
        opt._append_observation(P[i], Y[i], write=False)
    opt.fit_gp_model(); opt._hyper = None; opt._hyper_fits = 0            # allocate workspaces, then forget the optimum: cold again
    gp, res["cold_fit_gp_model_16_restarts_ms"] = t(opt.fit_gp_model)
    gp, res["warm_fit_gp_model_ms"] = t(opt.fit_gp_model)
    gp, res["warm_fit_gp_model_2_ms"] = t(opt.fit_gp_model)
    x, res["suggest_q1_pool1e4_ms"] = t(lambda: opt.suggest(1, gp))
    gp = opt.fit_gp_model()
    x, res["suggest_q16_believer_pool1e4_ms"] = t(lambda: opt.suggest(16, gp))
    gp = opt.fit_gp_model()
    x, res["suggest_q1000_topk_fps_pool1e4_ms"] = t(lambda: opt.suggest(1000, gp))
    opt.config.candidates_pool_size = 1_000_000
    gp = opt.fit_gp_model()
    x, res["suggest_q1000_topk_fps_pool1e6_ms"] = t(lambda: opt.suggest(1000, gp))
    assert x.shape == (1000, 5)
    out[f"n{n}"] = {k: round(v, 2) for k, v in res.items()}
    opt.close()
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/driver_timings.json", "w"), indent=1)
