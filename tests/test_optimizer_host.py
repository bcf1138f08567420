"""CPU tier: host-side logic of the BayesianOptimizer drop-in (constructor surface, CSV / resume contract,
LHS init, batch loop, failure conventions) with the oracle-backed engine double, and the sharding helpers."""
import inspect
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle_engine import OracleEngine

from bayesianoptimizer_b200.optimizer import BayesianOptimizer, GPConfig
from bayesianoptimizer_b200.simulators import DEFAULT_BOUNDS, CachedCSVSimulator


def _sim():
    z = np.load(os.path.join(GOLDEN_DIR, "csv_cache_rows.npz"))
    return CachedCSVSimulator(z["params"], z["outputs"])


def _cfg(**kw):
    base = dict(candidates_pool_size=512, num_restarts=3, refine_iters=5, hyper_restarts=2, hyper_maxiter=3, seed=0)
    base.update(kw)
    return GPConfig(**base)


def _opt(tmp_path, **kw):
    args = dict(simulator=_sim(), bounds_list=DEFAULT_BOUNDS, output_dir=str(tmp_path), n_initial_points=12, n_batches=2,
                batch_size=2, gp_config=_cfg(), engine_factory=OracleEngine, device=torch.device("cpu"))
    args.update(kw)
    return BayesianOptimizer(**args)


def test_constructor_is_superset_of_reference_signatures():
    """Bayesian7.py:202-218 keyword set (the call at scripts/run_optimization.py:116-127) + Bayesian.py:24 positional order."""
    params = list(inspect.signature(BayesianOptimizer.__init__).parameters)
    assert params[1:7] == ["simulator", "bounds_list", "output_dir", "n_initial_points", "n_batches", "batch_size"]
    for name in ("num_outputs", "svgp_threshold", "resume", "target_total", "device", "gp_config", "test_csv_path", "kwargs"):
        assert name in params
    for m in ("optimize", "fit_gp_model", "optimize_acquisition_function", "collect_initial_points", "run_simulation",
              "return_best_result", "_save_iteration_data", "suggest", "register", "maximize"):
        assert callable(getattr(BayesianOptimizer, m))


def test_bayesian_py_loop_and_csv_contract(tmp_path):
    opt = _opt(tmp_path)
    assert opt.bounds.shape == (2, 5) and opt.bounds.dtype == torch.float64           # Bayesian.py:42
    best_params, best_disp = opt.optimize()
    n_total = 12 + 2 * 2
    assert opt.train_X.shape == (n_total, 5) and opt.train_Y.shape == (n_total, 1)
    assert float(opt.train_X.min()) >= 0.0 and float(opt.train_X.max()) <= 1.0
    lines = open(opt.results_file).read().strip().split("\n")
    assert lines[0] == "n,eta,sigma_y,width,height,x_01,x_02,x_03,x_04,x_05,x_06,x_07,x_08"    # Bayesian7.py:269
    assert len(lines) == 1 + n_total                                                           # run_optimization.py:26-29
    row = np.array(lines[1].split(","), dtype=np.float64)
    assert row.shape == (13,) and len(lines[1].split(",")[0].split(".")[1]) == 16               # %.16f, Bayesian.py:66
    lo, hi = np.array(DEFAULT_BOUNDS).T
    assert np.all(row[:5] >= lo - 1e-12) and np.all(row[:5] <= hi + 1e-12)                      # physical units
    # objective = mean of the 8 displacements, maximised (Bayesian.py:140,98)
    objs = np.array([np.mean(d) for d in opt.displacements_list])
    np.testing.assert_allclose(opt.train_Y.reshape(-1).numpy(), objs, rtol=1e-12)
    assert np.allclose(best_params, opt.original_X[int(objs.argmax())])
    assert len(best_disp) == 8


def test_resume_and_target_total_semantics(tmp_path):
    opt = _opt(tmp_path, n_batches=0)
    opt.optimize()
    assert opt.train_X.shape[0] == 12
    # resume: rows are reloaded, no new LHS, loop runs until target_total (Bayesian7.py:271-289,635)
    opt2 = _opt(tmp_path, n_initial_points=0, n_batches=3, batch_size=2, resume=True, target_total=17,
                svgp_threshold=3000, test_csv_path="validation_set.csv")
    assert opt2.train_X.shape[0] == 12
    np.testing.assert_allclose(opt2.train_X.numpy(), opt.train_X.numpy(), atol=1e-12)
    best_params, best_value = opt2.optimize()
    assert opt2.train_X.shape[0] == 17 and isinstance(best_value, float)
    assert len(open(opt2.results_file).read().strip().split("\n")) == 18
    # without resume the file is truncated (Bayesian.py:56-60)
    opt3 = _opt(tmp_path, n_initial_points=1, n_batches=0)
    assert opt3.train_X.shape[0] == 0 and len(open(opt3.results_file).read().strip().split("\n")) == 1


def test_resume_tolerates_corrupt_rows_and_legacy_headers(tmp_path):
    """results/optimization_results2.csv:2386 has a stray space inside a number; two legacy files use disp_* headers."""
    p = tmp_path / "optimization_results.csv"
    hdr = "n,eta,sigma_y,width,height," + ",".join(f"disp_{i}" for i in range(1, 9))
    good = "0.5,100.0,200.0,3.0,4.0," + ",".join(["1.5"] * 8)
    bad = "0.5,100.0,200.067 7064061164856,3.0,4.0," + ",".join(["1.5"] * 8)
    p.write_text("\n".join([hdr, good, bad, good]) + "\n")
    opt = _opt(tmp_path, resume=True, n_initial_points=0, n_batches=0)
    assert opt.train_X.shape[0] == 2
    np.testing.assert_allclose(opt.train_Y.numpy().reshape(-1), [1.5, 1.5])


def test_simulator_failure_conventions(tmp_path):
    class Flaky:
        def __init__(self): self.k = 0
        def configure_geometry(self, w, h):
            if not 2.0 <= w <= 7.0: raise ValueError("Width must be between 2.0 and 7.0")
        def run_simulation(self, n, eta, s):
            self.k += 1
            return [None, np.array([1.0, 2.0, 3.0], dtype=np.float32), np.full(8, np.nan), np.arange(10.0)][self.k % 4]
        def cleanup(self): pass
    opt = _opt(tmp_path, simulator=Flaky())
    outs = [opt.run_simulation(np.array([0.5, 1.0, 1.0, 3.0, 3.0])) for _ in range(4)]
    assert np.array_equal(outs[0], [1, 2, 3, 0, 0, 0, 0, 0])          # short output padded (Bayesian.py:84-85)
    assert np.array_equal(outs[1], np.zeros(8))                         # NaN -> zeros (Bayesian.py:81-82)
    assert np.array_equal(outs[2], np.arange(8.0))                      # truncated to 8
    assert np.array_equal(outs[3], np.zeros(8))                         # None -> zeros
    assert np.array_equal(opt.run_simulation(np.array([0.5, 1.0, 1.0, 9.0, 3.0])), np.zeros(8))   # ValueError -> zeros


def test_cholesky_failure_retries_with_jitter(tmp_path):
    """Duplicate rows + ~zero noise: fit fails, the class retries with jitter 1e-2 (Bayesian6.py:482-488)."""
    opt = _opt(tmp_path, gp_config=_cfg(fit_hyperparameters=False, noise=0.0, kernel="rbf", lengthscale=[2.0] * 5))
    z = np.load(os.path.join(GOLDEN_DIR, "csv_cache_rows.npz"))
    for i in list(range(60)) + [12, 17]:
        opt._append_observation(z["params"][i], z["outputs"][i], write=False)
    gp = opt.fit_gp_model()
    assert gp.engine.fits == 1 and gp.engine.gp.L.shape[0] == 62      # first attempt raised, second (jitter) succeeded
    post = gp.posterior(opt.train_X[:3])
    assert post.mean.shape == (3, 1) and post.variance.shape == (3, 1)


def test_large_batch_uses_topk_then_fps(tmp_path):
    opt = _opt(tmp_path, batch_size=40, n_batches=1, gp_config=_cfg(believer_max_q=4, fit_hyperparameters=False))
    for x in opt.collect_initial_points():
        opt.register(x)
    batch = opt.optimize_acquisition_function(opt.fit_gp_model())
    assert batch.shape == (40, 5) and batch.dtype == torch.float64
    assert len({tuple(np.round(b, 12)) for b in batch.numpy()}) == 40     # FPS picks distinct points


def test_min_mode_flips_the_model_sign(tmp_path):
    opt = _opt(tmp_path, n_batches=1, objective_mode="min", objective_index=0)
    bp, bd = opt.optimize()
    objs = np.array([d[0] for d in opt.displacements_list])
    assert np.allclose(bp, opt.original_X[int(objs.argmin())])


def test_cached_simulator_duck_type():
    sim = _sim()
    with pytest.raises(ValueError):
        sim.configure_geometry(1.0, 3.0)                               # simulation/taichi.py:35-38
    sim.configure_geometry(3.0, 4.0)
    with pytest.raises(ValueError):
        sim.run_simulation(0.1, 1.0, 1.0)                              # simulation/taichi.py:64-71
    out = sim.run_simulation(0.5, 100.0, 200.0)
    assert out.dtype == np.float32 and out.shape == (8,)
    sim.cleanup()
    assert sim.cleaned


def test_lockstep_lbfgs_matches_scipy_lbfgsb():
    """hyperfit.lbfgs_lockstep (one batched LML call per step over all restarts) reaches SciPy L-BFGS-B's optimum
    of the same MAP objective, including an active bound (fit_gpytorch_mll stand-in, Bayesian.py:93)."""
    import scipy.optimize as so
    from bayesianoptimizer_b200.hyperfit import fit_map, log_prior_and_grad
    from conftest import synth_problem
    from oracle import gp_oracle as o
    rng = np.random.default_rng(0)
    for n, d, prior in ((150, 3, "lognormal"), (120, 4, "gamma"), (100, 2, None)):
        X, y = synth_problem(n, d, 1, 2)
        lo = np.log(np.array([0.025] * d + [1e-2, 1e-4])); hi = np.log(np.array([20.0] * d + [1e2, 1.0]))
        th0 = np.vstack([np.log([0.5] * d + [1.0, 1e-2]), rng.uniform(np.log(0.1), np.log(3), (3, d + 2))])
        th0[1:, d + 1] = np.log(1e-2)
        best, F, ths, Fs, nev = fit_map(OracleEngine(), X, y, "matern52", th0, lo, hi, prior=prior, maxiter=60)

        def negF(t):
            l, g = o.lml_and_grad(X, y, 0, np.exp(t[:d]), np.exp(t[d]), np.exp(t[d + 1]))
            lp, lg = log_prior_and_grad(t, d, prior)
            return -(l + lp[0]), -(g + lg[0])
        ref = max(-so.minimize(negF, t, jac=True, method="L-BFGS-B", bounds=list(zip(lo, hi)), options={"maxiter": 200}).fun
                  for t in th0)
        assert F >= ref - 1e-5 * max(1.0, abs(ref)), (n, d, prior, F, ref)
        assert np.all(best >= lo - 1e-12) and np.all(best <= hi + 1e-12)


def test_prior_gradients_against_finite_differences():
    from bayesianoptimizer_b200.hyperfit import log_prior_and_grad
    th = np.random.default_rng(1).standard_normal((3, 7)) * 0.5
    for prior in ("lognormal", "gamma", None):
        lp, g = log_prior_and_grad(th, 5, prior)
        fd = np.zeros_like(th)
        for k in range(7):
            e = np.zeros(7); e[k] = 1e-6
            fd[:, k] = (log_prior_and_grad(th + e, 5, prior)[0] - log_prior_and_grad(th - e, 5, prior)[0]) / 2e-6
        np.testing.assert_allclose(g, fd, atol=1e-6)


def test_linear_matern_kernel_loop_with_linear_variance_fit(tmp_path):
    """N4: GPConfig(kernel="linear_matern52") carries a fourth hyper-parameter (the LinearKernel variance of
    Bayesian6.py:471-473) through the lock-step fit, and suggestions stay pool-based for that kind."""
    opt = _opt(tmp_path, gp_config=_cfg(kernel="linear_matern52", hyper_restarts=3, hyper_maxiter=4), n_batches=1)
    opt.optimize()
    ls, s2, noise, lv = opt._hyper
    assert ls.shape == (5,) and lv > 0 and s2 > 0 and noise >= 1e-4
    assert opt._engine.gp.kind == 2 and opt._engine.gp.linear_variance == pytest.approx(lv)
    assert opt.train_X.shape == (14, 5)


def test_log_standardize_matches_oracle_restatement():
    from bayesianoptimizer_b200.transforms import LogStandardize
    from oracle import gp_oracle as o
    Y = np.random.default_rng(4).random((40, 8)) * 7 - 0.5
    tr, ref = LogStandardize.fit(torch.from_numpy(Y)), o.LogStandardize.fit(Y)
    assert tr.shift == pytest.approx(ref.shift, rel=1e-15)
    np.testing.assert_allclose(tr.forward(torch.from_numpy(Y)).numpy(), ref.forward(Y), rtol=1e-12, atol=1e-13)
    m, v = np.random.default_rng(5).standard_normal((30, 8)), np.random.default_rng(6).random(30)
    np.testing.assert_allclose(tr.inverse_mean(torch.from_numpy(m), torch.from_numpy(v)).numpy(),
                               ref.inverse_mean(m, v[:, None]), rtol=1e-12)


# ---- N2: batched SVGP predictor, host logic with the oracle-backed engine --------------------------------------
def _svgp_states(T=3, M=40, d=5, seed=0):
    from bayesianoptimizer_b200.svgp import SVGPTaskState
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(T):
        Ls = np.tril(rng.standard_normal((M, M)) * 0.05) + np.diag(0.3 + 0.5 * rng.random(M))
        out.append(SVGPTaskState(torch.from_numpy(rng.standard_normal((M, d))), torch.from_numpy(rng.standard_normal(M)),
                                 torch.from_numpy(Ls), torch.from_numpy(rng.uniform(0.5, 1.5, d)), float(rng.uniform(0.5, 2)),
                                 float(rng.uniform(0.05, 0.5)), float(rng.standard_normal()), float(rng.uniform(1e-4, 1e-2))))
    return out


def test_batch_svgp_predictor_host_logic():
    from bayesianoptimizer_b200.svgp import BatchSVGPPredictor
    from oracle import gp_oracle as o
    tasks = _svgp_states()
    bounds = torch.tensor(DEFAULT_BOUNDS, dtype=torch.float64).t()
    xlm, xls = torch.full((1, 5), 0.3, dtype=torch.float64), torch.full((1, 5), 1.7, dtype=torch.float64)
    pred = BatchSVGPPredictor(torch.device("cpu"), tasks, jitter=1e-6, bounds=bounds, x_log_mean=xlm, x_log_std=xls,
                              engine_factory=OracleEngine)
    U = np.random.default_rng(1).random((6000, 5))
    mean, var = pred.predict(U[:50])
    assert mean.shape == (3, 50) and var.shape == (3, 50)
    xs = o.svgp_transform_inputs(U, bounds.numpy(), xlm.numpy(), xls.numpy())
    ot = [o.SVGPTask(t.Z.numpy(), o.KERNEL_LINEAR_MATERN52, t.lengthscale.numpy(), t.outputscale, t.linear_variance, t.mean,
                     t.noise, 1e-6, t.var_mean.numpy(), t.var_chol.numpy()) for t in tasks]
    score = o.svgp_variance_score(ot, xs)
    np.testing.assert_allclose(pred.variance_score(U).numpy(), score, rtol=1e-12)
    pts, idx = pred.select_batch(U, 25, fps_start=0)
    big = np.argsort(-score, kind="stable")[:5000]
    assert set(idx.tolist()) <= set(big.tolist()) and pts.shape == (25, 5) and len(set(idx.tolist())) == 25
    ref_sel = o.fps(U[torch.topk(torch.from_numpy(score), 5000).indices.numpy()], 25, 0)
    assert idx.tolist() == torch.topk(torch.from_numpy(score), 5000).indices[ref_sel].tolist()


def test_svgp_tasks_from_state_dict():
    """Checkpoint layout written at Bayesian7.py:708-710 (gpytorch parameter names, softplus / GreaterThan constraints)."""
    import torch.nn.functional as F
    from bayesianoptimizer_b200.svgp import tasks_from_state_dict
    T, M, d = 8, 16, 5
    g = torch.Generator().manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g)
    msd = {"variational_strategy.inducing_points": r(T, M, d), "variational_strategy.variational_params_initialized": torch.tensor(1),
           "variational_strategy._variational_distribution.variational_mean": r(T, M),
           "variational_strategy._variational_distribution.chol_variational_covar": r(T, M, M),
           "mean_module.raw_constant": r(T), "covar_module.raw_outputscale": r(T),
           "covar_module.base_kernel.kernels.0.raw_variance": r(T, 1, 1), "covar_module.base_kernel.kernels.1.raw_lengthscale": r(T, 1, d)}
    lsd = {"noise_covar.raw_noise": r(T, 1)}
    tasks = tasks_from_state_dict(msd, lsd)
    assert len(tasks) == T and tasks[3].Z.shape == (M, d) and tasks[3].var_chol.shape == (M, M)
    assert tasks[3].outputscale == pytest.approx(float(F.softplus(msd["covar_module.raw_outputscale"][3].double())))
    assert tasks[5].noise == pytest.approx(float(F.softplus(lsd["noise_covar.raw_noise"][5, 0].double())) + 1e-4)
    np.testing.assert_allclose(tasks[2].lengthscale.numpy(), F.softplus(msd["covar_module.base_kernel.kernels.1.raw_lengthscale"][2, 0].double()).numpy())
    assert isinstance(tasks[2].linear_variance, float)
    # LinearKernel(ard_num_dims=d, batch_shape=[T]) (Bayesian7.py:162-166): current gpytorch registers raw_variance as (T, 1, d)
    msd["covar_module.base_kernel.kernels.0.raw_variance"] = r(T, 1, d)
    tasks = tasks_from_state_dict(msd, lsd)
    assert tasks[4].linear_variance.shape == (d,)
    np.testing.assert_allclose(tasks[4].linear_variance.numpy(),
                               F.softplus(msd["covar_module.base_kernel.kernels.0.raw_variance"][4, 0].double()).numpy())
    msd["covar_module.base_kernel.kernels.0.raw_variance"] = r(T, 2, d)
    with pytest.raises(ValueError):
        tasks_from_state_dict(msd, lsd)


def test_warm_refits_use_one_start_and_periodic_full_multistart(tmp_path):
    """Hyper-parameter refit policy: first fit = screened multi-start; warm refits refine the previous optimum only
    (the reference refits from a single start, Bayesian.py:92-93); every hyper_full_every-th fit is a full one again."""
    calls = []

    class CountingEngine(OracleEngine):
        def lml_grad_batched(self, X, y, thetas, kernel="matern52", mean=0.0):
            calls.append(np.asarray(thetas).reshape(-1, X.shape[1] + 2).shape[0])
            return super().lml_grad_batched(X, y, thetas, kernel, mean)

    opt = _opt(tmp_path, engine_factory=CountingEngine, gp_config=_cfg(hyper_restarts=5, hyper_refine=2, hyper_maxiter=2, hyper_full_every=3))
    for x in opt.collect_initial_points():
        opt.register(x)
    widths = []
    for _ in range(4):
        calls.clear()
        opt.fit_gp_model()
        widths.append(max(calls))
    assert widths[0] == 5            # cold: 5 restarts screened
    assert widths[1] == 1 and widths[2] == 1     # warm: the previous optimum only, no screening
    assert widths[3] == 5            # 4th fit = (fits - 1) % 3 == 0 -> full multi-start again


_REF_DRIVER = "/root/reference/scripts/run_optimization.py"


@pytest.mark.skipif(not os.path.exists(_REF_DRIVER), reason="the reference checkout is only mounted in the build container")
def test_reference_driver_runs_with_only_its_import_switched(tmp_path, monkeypatch):
    """INTEGRATION.md section 1, executed: the reference's OWN scripts/run_optimization.py (read from the mounted checkout, never
    copied) with line 4 switched to this package -- fresh run, then an automatic resume to a larger target.  The Taichi
    simulator module is stubbed by the cached-CSV double, the CUDA engine by the oracle-backed double (CPU tier)."""
    import sys
    import types
    import bayesianoptimizer_b200.optimizer as optmod
    src = open(_REF_DRIVER).read()
    line4 = "from optimization.Bayesian7 import BayesianOptimizer"
    assert src.splitlines()[3] == line4
    src = src.replace(line4, "from bayesianoptimizer_b200.optimizer import BayesianOptimizer")      # the one-line switch
    sims = []

    def make_sim(xml_path):
        assert xml_path == "config/setting.xml"
        sims.append(_sim())
        return sims[-1]

    fake = types.ModuleType("simulation.taichi")
    fake.MPMSimulator = make_sim
    monkeypatch.setitem(sys.modules, "simulation", types.ModuleType("simulation"))
    monkeypatch.setitem(sys.modules, "simulation.taichi", fake)
    monkeypatch.syspath_prepend("/root/reference")                      # config.config: the reference's own bounds
    monkeypatch.delitem(sys.modules, "config", raising=False)
    monkeypatch.delitem(sys.modules, "config.config", raising=False)
    monkeypatch.setattr(optmod, "GPEngine", lambda device: OracleEngine())
    monkeypatch.setattr(optmod, "GPConfig", lambda: _cfg(candidates_pool_size=256))
    ns = {"__name__": "run_optimization_under_test"}
    exec(compile(src, _REF_DRIVER, "exec"), ns)
    out = str(tmp_path / "results")
    best_params, best_value = ns["run_optimization"](total_evaluations=18, n_initial_points=12, batch_size=3, seed=0, output_dir=out)
    csv = os.path.join(out, "optimization_results.csv")
    assert ns["_count_existing_evals"](csv) == 18 and sims[0].cleaned and sims[0].calls == 18
    assert len(best_params) == 5 and np.isfinite(best_value)
    lo, hi = np.array(DEFAULT_BOUNDS).T
    assert np.all(best_params >= lo - 1e-9) and np.all(best_params <= hi + 1e-9)
    # second call: the driver detects the CSV, passes resume=True / target_total, and only the missing rows are added
    ns["run_optimization"](total_evaluations=22, n_initial_points=12, batch_size=3, seed=0, output_dir=out)
    assert ns["_count_existing_evals"](csv) == 22 and sims[1].calls == 4
    assert ns["run_optimization"](total_evaluations=22, n_initial_points=12, batch_size=3, seed=0, output_dir=out) == (None, None)
