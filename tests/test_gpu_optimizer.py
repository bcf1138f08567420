"""GPU tier: the BayesianOptimizer drop-in driving the real CUDA engine (BASELINE config 1: cached-CSV objective)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from oracle_engine import OracleEngine

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _sim():
    from bayesianoptimizer_b200.simulators import CachedCSVSimulator
    z = np.load(os.path.join(GOLDEN_DIR, "csv_cache_rows.npz"))
    return CachedCSVSimulator(z["params"], z["outputs"])


def test_suggestion_matches_oracle_backed_optimizer(tmp_path):
    """Same data, fixed hyper-parameters, pool sweep only: the CUDA engine and the oracle pick the same batch
    (Kriging-believer q=3 -> three sweeps + two appends)."""
    from bayesianoptimizer_b200.optimizer import BayesianOptimizer, GPConfig
    from bayesianoptimizer_b200.simulators import DEFAULT_BOUNDS
    cfg = GPConfig(candidates_pool_size=10_000, num_restarts=4, refine_iters=0, fit_hyperparameters=False,
                   lengthscale=[0.5, 0.4, 0.6, 0.8, 0.7], outputscale=1.3, noise=1e-3, seed=3)
    z = np.load(os.path.join(GOLDEN_DIR, "csv_cache_rows.npz"))
    picks = []
    for factory, dev, sub in ((None, torch.device("cuda", 0), "gpu"), (OracleEngine, torch.device("cpu"), "cpu")):
        out = tmp_path / sub
        opt = BayesianOptimizer(_sim(), DEFAULT_BOUNDS, str(out), 0, 1, 3, gp_config=cfg, engine_factory=factory, device=dev)
        for i in range(200):
            opt._append_observation(z["params"][i], z["outputs"][i], write=False)
        gp = opt.fit_gp_model()
        picks.append(opt.optimize_acquisition_function(gp).cpu().numpy())
        post = gp.posterior(opt.train_X[:5])
        assert post.mean.shape == (5, 1)
        opt.close()
    assert np.array_equal(picks[0], picks[1])


def test_full_loop_with_hyperparameter_fit(tmp_path):
    from bayesianoptimizer_b200.optimizer import BayesianOptimizer, GPConfig
    from bayesianoptimizer_b200.simulators import DEFAULT_BOUNDS
    cfg = GPConfig(candidates_pool_size=10_000, num_restarts=8, refine_iters=30, hyper_restarts=8, hyper_maxiter=15, seed=1)
    sim = _sim()
    opt = BayesianOptimizer(simulator=sim, bounds_list=DEFAULT_BOUNDS, output_dir=str(tmp_path), n_initial_points=40,
                            n_batches=2, batch_size=2, svgp_threshold=3000, resume=False, target_total=None,
                            test_csv_path="validation_set.csv", gp_config=cfg)
    assert opt.device.type == "cuda"
    best_params, best_disp = opt.optimize()
    assert opt.train_X.shape == (44, 5) and sim.calls == 44
    assert len(open(opt.results_file).read().strip().split("\n")) == 45
    ls, s2, noise, _ = opt._hyper
    assert ls.shape == (5,) and np.all(ls > 0) and s2 > 0 and noise >= cfg.min_noise
    # the fitted hyper-parameters are at least as likely as the defaults
    eng = opt._engine
    y, _, _, _ = opt._model_targets()
    th_fit = np.log(np.concatenate([ls, [s2], [noise]]))
    th_def = np.log(np.concatenate([np.full(5, 0.5), [1.0], [1e-3]]))
    lml, _, st = eng.lml_grad_batched(opt.train_X, y, np.vstack([th_fit, th_def]), cfg.kernel)
    assert st.tolist() == [0, 0] and lml[0] >= lml[1] - 1e-6
    assert len(best_disp) == 8 and np.all(np.isfinite(best_params))
    opt.close()


def test_large_batches_believer_beyond_16_and_q1000(tmp_path):
    """The reference's real batch sizes (main.py:15 batch_size=1000; Bayesian.py:105-112 q=batch_size): a Kriging-believer batch of
    q = 40 (believer_max_q raised: 40 sweeps, 39 appends crossing no refit) picks the same points as the oracle-backed class, and a
    q = 1000 suggestion (one sweep -> top-K_big -> device FPS, the shape of Bayesian7.py:676-688) returns 1000 distinct pool points
    identical to the oracle-backed selection."""
    from bayesianoptimizer_b200.optimizer import BayesianOptimizer, GPConfig
    from bayesianoptimizer_b200.simulators import DEFAULT_BOUNDS
    z = np.load(os.path.join(GOLDEN_DIR, "csv_cache_rows.npz"))
    got = {}
    for factory, dev, sub in ((None, torch.device("cuda", 0), "gpu"), (OracleEngine, torch.device("cpu"), "cpu")):
        cfg = GPConfig(candidates_pool_size=6_000, num_restarts=4, refine_iters=0, fit_hyperparameters=False,
                       lengthscale=[0.5, 0.4, 0.6, 0.8, 0.7], outputscale=1.3, noise=1e-3, seed=5, believer_max_q=40)
        opt = BayesianOptimizer(_sim(), DEFAULT_BOUNDS, str(tmp_path / sub), 0, 1, 40, gp_config=cfg, engine_factory=factory, device=dev)
        for i in range(150):
            opt._append_observation(z["params"][i], z["outputs"][i], write=False)
        gp = opt.fit_gp_model()
        kb = opt.suggest(40, gp).cpu().numpy()
        assert kb.shape == (40, 5) and len(np.unique(kb.round(12), axis=0)) == 40
        gp = opt.fit_gp_model()                                  # drop the 39 believer rows again
        big = opt.suggest(1000, gp).cpu().numpy()
        assert big.shape == (1000, 5) and len(np.unique(big, axis=0)) == 1000
        got[sub] = (kb, big)
        opt.close()
    assert np.array_equal(got["gpu"][0], got["cpu"][0])
    assert np.array_equal(got["gpu"][1], got["cpu"][1])
