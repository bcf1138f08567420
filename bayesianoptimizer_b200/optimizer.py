"""``BayesianOptimizer`` -- the reference's optimizer class surface on top of the B200 engine.

Source-compatible with billbearhunter/BayesianOptimizer so that ``scripts/run_optimization.py`` works by
changing only its import (line 4), and ``main.py``, the Taichi objective and the CSV writers plug in untouched:

* constructor: the superset signature of optimization/Bayesian7.py:202-218 (which covers Bayesian.py:24);
* methods of optimization/Bayesian.py: ``collect_initial_points`` (:68-73), ``run_simulation`` (:75-87),
  ``fit_gp_model`` (:89-94), ``optimize_acquisition_function`` (:96-113), ``return_best_result`` (:115-126),
  ``optimize`` (:128-180), ``_save_iteration_data`` (:62-66); attributes ``bounds``, ``train_X``, ``train_Y``,
  ``results_file``, ``device``;
* CSV / resume contract of optimization/Bayesian7.py:268-293 and scripts/run_optimization.py:21-31;
* ``suggest`` / ``register`` / ``maximize`` aliases named by BASELINE.json's north_star.

What changes is only the L2 body: SingleTaskGP + fit_gpytorch_mll + qLogEI + optimize_acqf
become GPEngine.fit / lml_grad_batched / sweep / refine / append (hand-written CUDA behind the C ABI).
PyTorch and NumPy here are plumbing: tensors, the lock-step L-BFGS driver of the hyper-parameter fit (hyperfit.py), CSV I/O.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence

import numpy as np
import torch

from .engine import GPEngine, NotPositiveDefiniteError
from .sobol import sobol_state

CSV_PARAM_COLS = ["n", "eta", "sigma_y", "width", "height"]


def _process_group():
    """(dist module, rank, world) of an initialised multi-rank process group, else (None, RANK from the launcher's
    environment, 1): under torchrun every rank runs the same driver script, so everything random or stateful below has to be
    decided once (rank 0) and shared."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist, dist.get_rank(), dist.get_world_size()
    return None, int(os.environ.get("RANK", "0") or 0), 1


@dataclass
class GPConfig:
    """Knobs of the surrogate / acquisition path (passed as ``gp_config=``, the kwarg of Bayesian7.py:215)."""
    kernel: str = "matern52"             # "matern52" | "rbf" (botorch >= 0.12 default) | "linear_matern52" (Bayesian6.py:471-473)
    linear_variance: float = 1.0         # LinearKernel variance (initial value when hyper-parameters are fitted)
    acquisition: str = "logei"           # "logei" (Bayesian.py:101) | "ei" | "ucb" | "var" (Bayesian7.py:670-671)
    beta: float = 2.0                    # UCB exploration weight
    candidates_pool_size: int = 1_000_000  # Sobol pool scored per suggestion (reference: 1024 raw / 10^4 LHS)
    sweep_mode: str = "auto"             # variance contraction of the pool sweep (bo_set_sweep_mode): "auto" | "fp64" | "i8x7" | "i8x8"
    num_restarts: int = 10               # starts refined per suggestion (Bayesian.py:108)
    refine_iters: int = 200              # refinement iterations (maxiter, Bayesian.py:111)
    believer_max_q: int = 16             # q <= this: Kriging-believer batch; larger q: top-K -> FPS (Bayesian7.py:676-688)
    K_BIG_CAP: int = 8000                # Bayesian7.py:66
    fit_hyperparameters: bool = True     # maximise the exact LML (fit_gpytorch_mll, Bayesian.py:93)
    hyper_restarts: int = 16             # batched random restarts screened by bo_lml_grad_batched
    hyper_refine: int = 4                # best screened restarts refined in lock step (one batched LML call per step)
    hyper_refine_warm: int = 1           # ... once a previous fit warm-starts the search (the reference refits from one start)
    hyper_full_every: int = 8            # every k-th refit repeats the full screened multi-start (guards against a stale local optimum)
    hyper_maxiter: int = 50              # lock-step L-BFGS iterations
    hyper_warm_curvature: bool = True    # warm refits also inherit the previous refit's L-BFGS curvature pairs
    hyper_prior: Optional[str] = "auto"  # "auto": botorch defaults (rbf -> lognormal, matern52 -> gamma); None = max. likelihood
    lengthscale: Optional[Sequence[float]] = None   # fixed / initial ARD lengthscales (unit cube)
    outputscale: float = 1.0
    noise: float = 1e-3
    min_noise: float = 1e-4              # botorch MIN_INFERRED_NOISE_LEVEL
    cholesky_jitter: float = 1e-2        # retry value of Bayesian6.py:482-488
    standardize: bool = True             # botorch Standardize outcome transform
    seed: Optional[int] = None           # pool / LHS seed (the reference never seeds, run_optimization.py:38)
    max_points: int = 49152              # exact-GP capacity guard (n = 32768 fits in 1 s / 25 GB on one B200); beyond it the
                                         # model is fitted on the incumbent + a random subset (the reference switches to SVGP)


class _Posterior:
    def __init__(self, mean, variance):
        self.mean, self.variance = mean, variance


class _Covar:
    def __init__(self, lengthscale):
        self.lengthscale = lengthscale


class GPModel:
    """What ``fit_gp_model`` returns: the model-object seam of SURVEY.md section 8b
    (``model.posterior(X).mean/.variance``, ``gp.covar_module.lengthscale`` printed at Bayesian.py:156)."""

    def __init__(self, engine, lengthscale, outputscale, noise, y_mean, y_std, sign):
        self.engine = engine
        self.covar_module = _Covar(torch.as_tensor(np.asarray(lengthscale, dtype=np.float64)).reshape(1, -1))
        self.outputscale, self.noise = float(outputscale), float(noise)
        self.y_mean, self.y_std, self.sign = float(y_mean), float(y_std), float(sign)

    def posterior(self, X):
        X = torch.as_tensor(X, dtype=torch.float64)
        shape = X.shape[:-1]
        mu, var = self.engine.posterior(X.reshape(-1, X.shape[-1]))
        mean = self.sign * (mu * self.y_std + self.y_mean)            # un-standardise (and undo the min->max flip)
        variance = var * self.y_std ** 2
        return _Posterior(mean.reshape(*shape, 1), variance.reshape(*shape, 1))


class BayesianOptimizer:
    def __init__(self, simulator, bounds_list: Sequence[Sequence[float]], output_dir: str, n_initial_points: int,
                 n_batches: int, batch_size: int, num_outputs: int = 8, svgp_threshold: int = 100, resume: bool = False,
                 target_total: Optional[int] = None, device: Optional[torch.device] = None,
                 gp_config: Optional[GPConfig] = None, test_csv_path: Optional[str] = None,
                 engine_factory: Optional[Callable] = None, **kwargs):
        if device is None:
            device = torch.device("cuda" if torch.cuda.is_available() else "cpu")       # Bayesian.py:28-29
        self.device = torch.device(device)
        self.gp_device = self.device                                                     # Bayesian7.py:219 name
        self.config = gp_config or GPConfig()
        self.simulator = simulator
        self.output_dir = output_dir
        os.makedirs(output_dir, exist_ok=True)
        self.n_initial_points = int(n_initial_points)
        self.n_batches = int(n_batches)
        self.batch_size = int(batch_size)
        self.num_outputs = int(num_outputs)
        self.svgp_threshold = svgp_threshold          # accepted for compatibility; the exact GP scales to n = 8192+
        self.resume = bool(resume)
        self.target_total = target_total
        self.test_csv_path = test_csv_path            # accepted; validation metrics are outside the hot path
        # objective: mean of the outputs, maximised (Bayesian.py:140,98); Bayesian7's kwargs are honoured
        self.objective_mode = str(kwargs.get("objective_mode", "max")).lower()
        self.objective_index = kwargs.get("objective_index", None)
        self.objective_weights = kwargs.get("objective_weights", None)

        self.bounds = torch.tensor([list(map(float, b)) for b in bounds_list], dtype=torch.float64, device=self.device).t()
        self.physical_bounds = self.bounds.t().cpu().numpy()                              # (d, 2), Bayesian7.py:225
        self.dim = int(self.bounds.shape[1])
        self.train_X = torch.empty((0, self.dim), dtype=torch.float64, device=self.device)
        self.train_Y = torch.empty((0, 1), dtype=torch.float64, device=self.device)
        self.original_X = []
        self.displacements_list = []

        self.results_file = os.path.join(output_dir, "optimization_results.csv")
        self.results_csv_path = self.results_file                                         # Bayesian7.py:255 name
        self._engine_factory = engine_factory or (lambda: GPEngine(self.device))
        self._engine = None
        self._hyper = None            # (lengthscale[d], outputscale, noise, linear variance) carried between refits (warm start)
        self._hyper_fits = 0          # hyper-parameter refits so far (every hyper_full_every-th one is a full multi-start)
        self._lbfgs_memory = {}       # curvature pairs carried between warm refits (hyperfit.lbfgs_lockstep)
        self._y_mean, self._y_std = 0.0, 1.0
        self._suggest_count = 0
        self._rng = np.random.default_rng(self.config.seed)
        self._replicas_synced = False
        self._init_results_file()

    # ------------------------------------------------------------------ CSV / resume -------------
    def _csv_header(self):
        return CSV_PARAM_COLS[:self.dim] + [f"x_{i:02d}" for i in range(1, self.num_outputs + 1)]

    def _init_results_file(self):
        """Bayesian.py:56-60 truncates; with resume=True an existing file is loaded instead (Bayesian7.py:271-289)."""
        if self.resume and os.path.exists(self.results_file):
            self._load_existing()
            return
        if _process_group()[1] != 0:
            return                                    # one writer: rank 0 owns optimization_results.csv
        with open(self.results_file, "w") as f:
            f.write(",".join(self._csv_header()) + "\n")

    def _sync_replicas(self):
        """Multi-GPU driver mode (torchrun): every rank runs this class on its own GPU and shards the sweeps and
        the restart screening, so all ranks must hold the same data, the same random stream and the same warm-start state.
        A seed drawn by rank 0 (GPConfig.seed = None seeds every process differently), its observations and hyper-parameters are
        broadcast once, the first time a process group is seen; afterwards every rank performs the same draws.  The
        simulator runs and the CSV is written on rank 0 only (``register``)."""
        dist, rank, world = _process_group()
        if dist is None or self._replicas_synced:
            return
        payload = [None]
        if rank == 0:
            payload[0] = {"rng": int(self._rng.integers(0, 2 ** 63 - 1)), "X": [np.asarray(x) for x in self.original_X],
                          "D": [np.asarray(v) for v in self.displacements_list], "hyper": self._hyper,
                          "hyper_fits": self._hyper_fits, "suggest_count": self._suggest_count}
        dist.broadcast_object_list(payload, src=0)
        st = payload[0]
        # every rank (0 included) restarts its generator from the one broadcast seed: SciPy's LHS spawns children from the
        # generator's seed sequence, which a copied bit-generator state would not carry
        self._rng = np.random.default_rng(st["rng"])
        if rank != 0:
            self.train_X = torch.empty((0, self.dim), dtype=torch.float64, device=self.device)
            self.train_Y = torch.empty((0, 1), dtype=torch.float64, device=self.device)
            self.original_X, self.displacements_list = [], []
            for x, v in zip(st["X"], st["D"]):
                self._append_observation(x, v, write=False)
            self._hyper, self._hyper_fits, self._suggest_count = st["hyper"], st["hyper_fits"], st["suggest_count"]
        self._replicas_synced = True

    def _load_existing(self):
        """Reload rows (physical units) and re-normalise; tolerates the legacy ``disp_*`` headers and drops
        malformed rows such as results/optimization_results2.csv:2386."""
        import pandas as pd
        try:
            df = pd.read_csv(self.results_file)
        except Exception as e:  # pragma: no cover
            print(f"[Resume] Failed to load CSV: {e}")
            return
        if df.empty:
            return
        df = df.apply(pd.to_numeric, errors="coerce").dropna()
        ncol = self.dim + self.num_outputs
        if df.shape[1] < ncol:
            print("[Resume] CSV has too few columns; ignoring it")
            return
        arr = df.to_numpy(dtype=np.float64)[:, :ncol]
        print(f"[Resume] Loaded {arr.shape[0]} samples.")
        for row in arr:
            self._append_observation(row[:self.dim], row[self.dim:], write=False)

    def _save_iteration_data(self, params_numpy, displacements_numpy):
        row = list(np.asarray(params_numpy, dtype=np.float64).tolist()) + list(np.asarray(displacements_numpy, dtype=np.float64).tolist())
        with open(self.results_file, "a") as f:
            f.write(",".join([f"{v:.16f}" for v in row]) + "\n")

    # ------------------------------------------------------------------ transforms ---------------
    def _normalize(self, x_phys):
        lo, hi = self.physical_bounds[:, 0], self.physical_bounds[:, 1]
        return (np.asarray(x_phys, dtype=np.float64) - lo) / (hi - lo)

    def _unnormalize(self, x_unit):
        lo, hi = self.physical_bounds[:, 0], self.physical_bounds[:, 1]
        return np.asarray(x_unit, dtype=np.float64) * (hi - lo) + lo

    def _objective(self, displacements):
        d = np.asarray(displacements, dtype=np.float64)
        if self.objective_weights is not None:
            return float(np.dot(d[:len(self.objective_weights)], np.asarray(self.objective_weights, dtype=np.float64)))
        if self.objective_index is not None:
            return float(d[int(self.objective_index)])
        return float(np.mean(d))                                                          # Bayesian.py:140

    def _append_observation(self, x_phys, displacements, write=True):
        """The register step (Bayesian.py:143-148): grow train_X / train_Y, remember originals, append the CSV row."""
        x_phys = np.asarray(x_phys, dtype=np.float64)
        displacements = np.asarray(displacements, dtype=np.float64)
        x_unit = torch.as_tensor(self._normalize(x_phys), dtype=torch.float64, device=self.device)
        self.train_X = torch.cat([self.train_X, x_unit.unsqueeze(0)])
        self.train_Y = torch.cat([self.train_Y, torch.tensor([[self._objective(displacements)]], dtype=torch.float64,
                                                             device=self.device)])
        self.original_X.append(x_phys)
        self.displacements_list.append(displacements)
        if write:
            self._save_iteration_data(x_phys, displacements)

    # ------------------------------------------------------------------ reference surface --------
    def collect_initial_points(self):
        """Latin-hypercube points in [0,1]^d (Bayesian.py:68-73)."""
        from scipy.stats import qmc
        self._sync_replicas()
        print(f"Generating {self.n_initial_points} initial points using Latin Hypercube Sampling...")
        sampler = qmc.LatinHypercube(d=self.dim, seed=self._rng)
        pts = sampler.random(n=self.n_initial_points) if self.n_initial_points > 0 else np.empty((0, self.dim))
        return torch.from_numpy(pts).to(dtype=torch.float64, device=self.device)

    def run_simulation(self, params_numpy):
        """Objective plug (Bayesian.py:75-87): physical parameters -> 8 displacements; zeros on failure."""
        p = np.asarray(params_numpy, dtype=np.float64).reshape(-1)
        n, eta, sigma_y, width, height = (float(v) for v in p[:5])
        try:
            self.simulator.configure_geometry(width, height)
            disp = self.simulator.run_simulation(n, eta, sigma_y)
        except Exception as e:                                   # simulator failure -> skip-as-zeros (Bayesian7.py:339-352)
            print(f"[run_simulation] simulator failed: {e}")
            disp = None
        k = self.num_outputs
        if disp is None or len(disp) == 0 or np.isnan(np.asarray(disp, dtype=np.float64)).any():
            return np.zeros(k)
        disp = np.asarray(disp, dtype=np.float64).reshape(-1)
        if len(disp) < k:
            return np.concatenate([disp, np.zeros(k - len(disp))])
        return disp[:k].copy()

    def _model_targets(self):
        """Signed, standardised targets: maximisation of the (mean-displacement) objective."""
        y = self.train_Y.reshape(-1)
        sign = -1.0 if self.objective_mode == "min" else 1.0
        y = sign * y
        if self.config.standardize and y.numel() > 1:
            mu = float(y.mean())
            sd = float(y.std(unbiased=True))
            if not sd >= 1e-8:
                sd = 1.0
        else:
            mu, sd = 0.0, 1.0
        return (y - mu) / sd, mu, sd, sign

    def _engine_get(self):
        if self._engine is None:
            self._engine = self._engine_factory()
            if self.config.sweep_mode != "auto":
                self._engine.set_sweep_mode(self.config.sweep_mode)
        return self._engine

    def _fit_hyperparameters(self, eng, X, y):
        """Exact-MLL / MAP fit (fit_gpytorch_mll, Bayesian.py:93): R random restarts screened in one batched
        K7 call, the best few refined by a lock-step box-projected L-BFGS (hyperfit.py) whose every step is one
        batched LML+gradient evaluation on the device; warm-started from the previous refit."""
        from .hyperfit import fit_map, log_prior_and_grad
        cfg, d = self.config, self.dim
        prior = cfg.hyper_prior
        if prior == "auto":
            prior = "lognormal" if cfg.kernel == "rbf" else "gamma"
        lin = cfg.kernel in ("linear_matern52", "linear+matern52")
        lo = np.concatenate([np.full(d, math.log(0.025)), [math.log(1e-2)], [math.log(cfg.min_noise)]])
        hi = np.concatenate([np.full(d, math.log(20.0)), [math.log(1e2)], [math.log(1.0)]])
        if self._hyper is not None:
            ls0, s20, nz0, lv0 = self._hyper
        else:
            ls0 = np.full(d, 0.5) if cfg.lengthscale is None else np.broadcast_to(np.asarray(cfg.lengthscale, dtype=np.float64), (d,))
            s20, nz0, lv0 = cfg.outputscale, cfg.noise, cfg.linear_variance
        th0 = np.concatenate([np.log(ls0), [math.log(s20)], [math.log(max(nz0, cfg.min_noise))]])
        if lin:                                       # one more column: log LinearKernel variance
            lo, hi = np.append(lo, math.log(1e-4)), np.append(hi, math.log(1e2))
            th0 = np.append(th0, math.log(max(lv0, 1e-4)))
        th0 = np.clip(th0, lo, hi)
        self._hyper_fits += 1
        warm_only = (self._hyper is not None and int(cfg.hyper_refine_warm) <= 1 and
                     (int(cfg.hyper_full_every) <= 0 or (self._hyper_fits - 1) % int(cfg.hyper_full_every) != 0))
        R = 1 if warm_only else max(int(cfg.hyper_restarts), 1)      # warm refits skip the screening of random restarts
        thetas = np.tile(th0, (R, 1))
        if R > 1:
            thetas[1:, :d] = self._rng.uniform(math.log(0.1), math.log(3.0), size=(R - 1, d))
            thetas[1:, d] = self._rng.uniform(math.log(0.3), math.log(3.0), size=R - 1)
            thetas[1:, d + 1] = self._rng.uniform(math.log(cfg.min_noise), math.log(1e-1), size=R - 1)
            if lin:
                thetas[1:, d + 2] = self._rng.uniform(math.log(1e-2), math.log(1e1), size=R - 1)
        unpack = lambda t: (np.exp(t[:d]), float(np.exp(t[d])), float(np.exp(t[d + 1])), float(np.exp(t[d + 2])) if lin else 0.0)
        if R == 1:                                    # nothing to screen: refine the (warm) start directly
            # the L-BFGS curvature pairs survive from one warm refit to the next (the objective gains one observation per
            # iteration): the first step is a quasi-Newton step, not a short steepest-ascent one
            if self._lbfgs_memory.get("key") != (cfg.kernel, len(th0)):
                self._lbfgs_memory = {"key": (cfg.kernel, len(th0))}
            th, F, _, _, _ = fit_map(eng, X, y, cfg.kernel, thetas, lo, hi, prior=prior, maxiter=int(cfg.hyper_maxiter),
                                     memory=self._lbfgs_memory if cfg.hyper_warm_curvature else None)
            return unpack(th if np.isfinite(F) else th0)
        self._lbfgs_memory = {}                       # a full multi-start may land in another basin: its curvature starts afresh
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            from .dist import sharded_lml_grad                    # restarts are independent: shard the screening over the ranks
            lml, _, status = sharded_lml_grad(eng, X, y, thetas, cfg.kernel, 0.0, dist.get_rank(), dist.get_world_size())
        else:
            lml, _, status = eng.lml_grad_batched(X, y, thetas, cfg.kernel)
        score = np.asarray(lml, dtype=np.float64) + log_prior_and_grad(thetas, d, prior)[0]
        score = np.where(np.asarray(status) == 0, score, -np.inf)
        if not np.isfinite(score).any():
            return unpack(th0)
        n_refine = cfg.hyper_refine if (self._hyper is None or int(cfg.hyper_refine_warm) <= 1) else min(cfg.hyper_refine, cfg.hyper_refine_warm)
        keep = np.argsort(-score)[:max(1, int(n_refine))]
        if 0 not in keep and np.isfinite(score[0]):
            keep = np.concatenate([keep[:-1], [0]]) if len(keep) > 1 else np.array([0])      # always refine the warm start
        th, F, _, _, _ = fit_map(eng, X, y, cfg.kernel, thetas[keep], lo, hi, prior=prior, maxiter=int(cfg.hyper_maxiter))
        if not np.isfinite(F) or F < score[keep].max():
            th = thetas[keep[int(np.argmax(score[keep]))]]
        return unpack(th)

    def fit_gp_model(self):
        """Fit the exact GP on the normalised data (Bayesian.py:89-94) and return the model handle."""
        self._sync_replicas()
        if self.train_X.shape[0] == 0:
            raise RuntimeError("fit_gp_model: no observations")
        eng = self._engine_get()
        cfg, d = self.config, self.dim
        y, mu, sd, sign = self._model_targets()
        self._y_mean, self._y_std = mu, sd
        X = self.train_X.to(eng.device) if hasattr(eng, "device") else self.train_X
        y = y.to(X.device)
        if X.shape[0] > int(cfg.max_points):
            keep = self._rng.choice(X.shape[0], size=int(cfg.max_points), replace=False)
            keep[0] = int(torch.argmax(y).item())                     # never drop the incumbent
            keep = torch.as_tensor(np.unique(keep), device=X.device)
            X, y = X[keep], y[keep]
        if cfg.fit_hyperparameters and self.train_X.shape[0] >= 2 * d:
            ls, s2, noise, lv = self._fit_hyperparameters(eng, X, y)
        elif self._hyper is not None:
            ls, s2, noise, lv = self._hyper
        else:
            ls = np.full(d, 0.5) if cfg.lengthscale is None else np.broadcast_to(np.asarray(cfg.lengthscale, dtype=np.float64), (d,)).copy()
            s2, noise, lv = cfg.outputscale, cfg.noise, cfg.linear_variance
        self._hyper = (np.asarray(ls, dtype=np.float64), float(s2), float(noise), float(lv))
        try:
            eng.fit(X, y, cfg.kernel, ls, s2, noise, mean=0.0, jitter=0.0, linear_variance=lv)
        except NotPositiveDefiniteError as e:
            # retry-with-jitter convention of Bayesian6.py:482-488
            print(f"[fit_gp_model] Cholesky failed at pivot {e.pivot}; retrying with jitter {cfg.cholesky_jitter:g}")
            eng.fit(X, y, cfg.kernel, ls, s2, noise, mean=0.0, jitter=cfg.cholesky_jitter, linear_variance=lv)
        return GPModel(eng, ls, s2, noise, mu, sd, sign)

    def _best_f(self):
        """Incumbent in the model's (signed, standardised) space -- Bayesian.py:98 / Bayesian2.py:221-227."""
        y, _, _, _ = self._model_targets()
        return float(y.max())

    def _sweep_topk(self, eng, sob, best_f, k):
        """Pool scan; sharded over the ranks of an initialised process group, one (value, index) all-gather."""
        cfg = self.config
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            from .dist import sharded_sweep
            return sharded_sweep(eng, cfg.acquisition, best_f, cfg.beta, sob, cfg.candidates_pool_size, k,
                                 dist.get_rank(), dist.get_world_size())
        return eng.sweep(cfg.acquisition, best_f, cfg.beta, sobol=sob, count=cfg.candidates_pool_size, topk=k)

    def optimize_acquisition_function(self, gp):
        """Next batch of q = batch_size points in [0,1]^d (Bayesian.py:96-113): Sobol pool sweep -> top-k starts ->
        batched gradient refinement -> best point; q > 1 by Kriging-believer appends (or top-K -> FPS for large q)."""
        return self.suggest(self.batch_size, gp)

    def suggest(self, q: Optional[int] = None, gp: Optional[GPModel] = None):
        cfg = self.config
        self._sync_replicas()
        q = self.batch_size if q is None else int(q)
        if gp is None:
            gp = self.fit_gp_model()
        eng = gp.engine
        best_f = self._best_f()
        k = max(1, min(int(cfg.num_restarts), 64))
        if q > cfg.believer_max_q:
            return self._suggest_topk_fps(eng, best_f, q)
        out = []
        for j in range(q):
            seed = (0 if cfg.seed is None else int(cfg.seed)) * 1_000_003 + self._suggest_count
            self._suggest_count += 1
            sob = sobol_state(self.dim, seed)
            vals, idx = self._sweep_topk(eng, sob, best_f, k)
            keep = idx >= 0
            starts = eng.sobol_points(sob, idx[keep])
            x, v = starts, vals[keep]
            if cfg.refine_iters > 0:
                x, v = eng.refine(starts, cfg.acquisition, best_f, cfg.beta, iters=cfg.refine_iters)
            order = torch.argsort(v, descending=True).tolist()
            xb = x[order[0]]
            for cand in order:                           # best refined point that is not (numerically) already in the batch
                if all(float(torch.linalg.norm(x[cand].to(self.device) - p)) > 1e-9 for p in out):
                    xb = x[cand]
                    break
            out.append(xb.to(self.device))
            if j + 1 < q:
                try:
                    eng.append(xb)                       # Kriging believer: y = mu(x), alpha' = [alpha; 0]
                except NotPositiveDefiniteError:
                    # the pick sits on top of an observation (bordered matrix not PD): the model is unchanged, so the next
                    # sweep would return the same point -- condition on a copy nudged off the duplicate instead
                    nudged = torch.clamp(xb + 1e-4 * torch.as_tensor(self._rng.standard_normal(self.dim), dtype=xb.dtype,
                                                                     device=xb.device), 0.0, 1.0)
                    try:
                        eng.append(nudged)
                    except NotPositiveDefiniteError:
                        print("[suggest] Kriging-believer append failed twice at a duplicate point; the batch may repeat it")
        return torch.stack(out).to(dtype=torch.float64)

    def _suggest_topk_fps(self, eng, best_f, q):
        """Large batches: one sweep -> top K_big -> farthest-point sampling (the shape of Bayesian7.py:676-688)."""
        cfg = self.config
        seed = (0 if cfg.seed is None else int(cfg.seed)) * 1_000_003 + self._suggest_count
        self._suggest_count += 1
        sob = sobol_state(self.dim, seed)
        N = int(cfg.candidates_pool_size)
        _, _, _, _, acq = eng.sweep(cfg.acquisition, best_f, cfg.beta, sobol=sob, count=N, topk=1, return_all=True)
        K_big = int(min(max(5000, 20 * q), cfg.K_BIG_CAP, N))
        K_big = max(K_big, min(q, N))
        _, idx = eng.topk_scores(acq, K_big)
        pts = eng.sobol_points(sob, idx)
        sel = eng.fps(pts, min(q, K_big), 0)          # device FPS from the best-scoring candidate (Bayesian7.py:685)
        return pts[sel].to(self.device, dtype=torch.float64)

    def register(self, x_scaled, displacements=None):
        """Evaluate (if needed) and record one point: the torch.cat + CSV block of Bayesian.py:143-148."""
        self._sync_replicas()
        x_unit = torch.as_tensor(x_scaled, dtype=torch.float64).detach().cpu().numpy().reshape(-1)
        x_phys = self._unnormalize(x_unit)
        dist, rank, world = _process_group()
        if dist is not None:
            # one simulation, one CSV row: rank 0 evaluates and writes, the replicas receive the displacements
            box = [np.asarray(self.run_simulation(x_phys) if displacements is None else displacements, dtype=np.float64)
                   if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            displacements = box[0]
        elif displacements is None:
            displacements = self.run_simulation(x_phys)
        self._append_observation(x_phys, displacements, write=(rank == 0))
        return self._objective(displacements)

    def return_best_result(self):
        """Best parameters and their displacements (Bayesian.py:115-126)."""
        if self.train_Y.numel() == 0:
            return None, None
        obj = self.train_Y.reshape(-1)
        best_idx = int((torch.argmin(obj) if self.objective_mode == "min" else torch.argmax(obj)).item())
        best_params = self.original_X[best_idx]
        best_disp = self.displacements_list[best_idx]
        print("\n--- Best Result Found ---")
        print(f"Best Parameters (Original Scale): {best_params}")
        print(f"Best Displacements: {best_disp}")
        print(f"Best Objective: {obj[best_idx].item()}")
        return best_params, best_disp

    def optimize(self):
        """Batch optimisation loop (Bayesian.py:128-180); with ``target_total`` the loop runs until the global
        evaluation count is reached (Bayesian7.py:614-733, scripts/run_optimization.py:66-98)."""
        print("Starting optimization...")
        self._sync_replicas()
        n_have = self.train_X.shape[0]
        n_init = max(0, self.n_initial_points - n_have) if self.target_total is not None else self.n_initial_points
        if n_init > 0:
            keep = self.n_initial_points
            self.n_initial_points = n_init
            initial = self.collect_initial_points()
            self.n_initial_points = keep
            for x_scaled in initial:
                self.register(x_scaled)
        if self.target_total is not None:
            while self.train_X.shape[0] < int(self.target_total):
                q = min(self.batch_size, int(self.target_total) - self.train_X.shape[0])
                print(f"\n=== Iteration: {self.train_X.shape[0]} samples ===")
                gp = self.fit_gp_model()
                for x_scaled in self.suggest(q, gp):
                    self.register(x_scaled)
            if self.train_X.shape[0] == 0:
                return None, None
            best_params, _ = self.return_best_result()
            obj = self.train_Y.reshape(-1)
            best_value = float((obj.min() if self.objective_mode == "min" else obj.max()).item())
            print("\nOptimization completed!")
            return best_params, best_value
        for i in range(self.n_batches):
            print(f"\n--- Batch {i + 1}/{self.n_batches} ---")
            gp = self.fit_gp_model()
            print(f"Learned Lengthscale: {gp.covar_module.lengthscale.detach().cpu().numpy()}")
            batch = self.optimize_acquisition_function(gp)
            print(f"New Scaled Candidates:\n{batch.cpu().numpy()}")
            for x_scaled in batch:
                val = self.register(x_scaled)
                xo = self.original_X[-1]
                print(f"  Evaluated point (original scale): n={xo[0]:.3f}, eta={xo[1]:.3f}, ... | Avg Disp: {val:.4f}")
        best_params, best_disp = self.return_best_result()
        print("\nOptimization completed!")
        return best_params, best_disp

    maximize = optimize      # north_star alias

    def close(self):
        if self._engine is not None and hasattr(self._engine, "close"):
            self._engine.close()
            self._engine = None
