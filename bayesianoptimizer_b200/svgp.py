"""Predictive sweep of the batched sparse variational GP the shipped driver trains.

``optimization/Bayesian7.py`` fits ``BatchSVGP`` (T = 8 independent tasks over ``batch_shape``, M <= 2048 inducing points
each, kernel ScaleKernel(Linear + Matern-5/2), ConstantMean, GaussianLikelihood; :129-195) with an SGD/ELBO loop and then
scans a candidate pool chunk by chunk (:664-671), copies every score to the host, takes ``torch.topk`` there and runs a
CPU farthest-point sampling (:676-688).  Training stays where it is (out of scope); this module takes the TRAINED state --
inducing points, variational mean / Cholesky factor, kernel and likelihood parameters -- and evaluates the predictive
distribution with the same fused CUDA sweep as the exact path: one ``bo_svgp_load`` per task, then ``bo_posterior`` /
``bo_sweep`` (two triangular DMMA contractions per candidate block), the variance-sum score, top-K and FPS on the device.

PyTorch here is plumbing (tensors, the elementwise input transform, the T-way sum of the per-task scores).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

from .engine import MIN_VARIANCE, GPEngine


@dataclass
class SVGPTaskState:
    """Natural (constrained) parameters of one task."""
    Z: torch.Tensor                 # (M, d) inducing points in the model's (standardised-log) input space
    var_mean: torch.Tensor          # (M,) whitened variational mean
    var_chol: torch.Tensor          # (M, M) chol_variational_covar (lower triangle used)
    lengthscale: torch.Tensor       # (d,)
    outputscale: float
    linear_variance: object         # float, or a (d,) tensor: one LinearKernel variance per input dimension
    mean: float
    noise: float


def tasks_from_state_dict(model_sd: Dict[str, torch.Tensor], likelihood_sd: Dict[str, torch.Tensor],
                          noise_lower_bound: float = 1e-4) -> List[SVGPTaskState]:
    """Split the checkpoint ``{"model": gp_model.state_dict(), "likelihood": likelihood.state_dict()}`` written at
    optimization/Bayesian7.py:708-710 into per-task states.  Parameter names and constraints follow gpytorch
    [3P-recall, SURVEY.md App. A]: raw parameters go through softplus (``Positive``), the likelihood noise through
    ``GreaterThan(1e-4)`` (softplus + lower bound)."""
    g = lambda *names: next(model_sd[n] for n in names if n in model_sd)
    Z = g("variational_strategy.inducing_points").to(torch.float64)                                  # (T, M, d)
    m = g("variational_strategy._variational_distribution.variational_mean").to(torch.float64)       # (T, M)
    Ls = g("variational_strategy._variational_distribution.chol_variational_covar").to(torch.float64)
    T, M, d = Z.shape
    const = g("mean_module.raw_constant", "mean_module.constant").to(torch.float64).reshape(T)
    s2 = F.softplus(g("covar_module.raw_outputscale").to(torch.float64)).reshape(T)
    # LinearKernel(ard_num_dims=d, batch_shape=[T]) (Bayesian7.py:162-166): raw_variance is (T, 1, d) in current gpytorch -- one
    # variance per input dimension -- and (T, 1, 1) where ard_num_dims is not honoured; both are accepted
    raw_v = g("covar_module.base_kernel.kernels.0.raw_variance").to(torch.float64)
    if raw_v.numel() == T:
        v = F.softplus(raw_v).reshape(T)
    elif raw_v.numel() == T * d:
        v = F.softplus(raw_v).reshape(T, d)
    else:
        raise ValueError(f"LinearKernel raw_variance has shape {tuple(raw_v.shape)}; expected (T, 1, 1) or (T, 1, d) with T = {T}, d = {d}")
    ls = F.softplus(g("covar_module.base_kernel.kernels.1.raw_lengthscale").to(torch.float64)).reshape(T, d)
    raw_noise = likelihood_sd["noise_covar.raw_noise"].to(torch.float64).reshape(T)
    noise = F.softplus(raw_noise) + noise_lower_bound
    lin_v = lambda t: float(v[t]) if v.ndim == 1 else v[t].clone()
    return [SVGPTaskState(Z[t], m[t], Ls[t], ls[t], float(s2[t]), lin_v(t), float(const[t]), float(noise[t])) for t in range(T)]


class BatchSVGPPredictor:
    """T task handles on one device.  ``jitter`` is gpytorch's variational_cholesky_jitter: 1e-4 for the float32 model
    the reference trains (Bayesian7.py:220), 1e-6 for a float64 one; ``min_variance`` likewise 1e-3 / 1e-6."""

    def __init__(self, device, tasks: Sequence[SVGPTaskState], jitter: float = 1e-4, kernel: str = "linear_matern52",
                 bounds: Optional[torch.Tensor] = None, x_log_mean: Optional[torch.Tensor] = None,
                 x_log_std: Optional[torch.Tensor] = None, engine_factory=None):
        self.device = torch.device(device)
        self.tasks = list(tasks)
        self.jitter = float(jitter)
        self.engines = []
        make = engine_factory or (lambda: GPEngine(self.device))
        for t in self.tasks:
            eng = make()
            eng.load_svgp(t.Z, t.var_mean, t.var_chol, kernel, t.lengthscale, t.outputscale, t.linear_variance, t.mean,
                          t.noise, self.jitter)
            self.engines.append(eng)
        f64 = lambda a: None if a is None else torch.as_tensor(a, dtype=torch.float64, device=self.device)
        self.bounds, self.x_log_mean, self.x_log_std = f64(bounds), f64(x_log_mean), f64(x_log_std)
        # one stream per task: a 10^4-candidate pool is only 79 blocks of 128 -- less than one wave of the 148 SMs -- so the
        # T task sweeps run concurrently and the block scheduler packs their CTAs onto the free SMs
        self._streams = [torch.cuda.Stream(device=self.device) for _ in self.engines] if self.device.type == "cuda" else None
        self._pool = None

    @property
    def num_tasks(self):
        return len(self.engines)

    def transform_inputs(self, x_unit):
        """Unit cube -> physical -> log -> standardised (BatchSVGP._transform_inputs, Bayesian7.py:178-188)."""
        x = torch.as_tensor(x_unit, dtype=torch.float64, device=self.device)
        if self.bounds is None:
            return x
        lo, hi = self.bounds[0], self.bounds[1]
        x_log = torch.log((x * (hi - lo) + lo).clamp(min=1e-6))
        return (x_log - self.x_log_mean) / self.x_log_std

    def predict(self, x_unit, min_variance: float = MIN_VARIANCE):
        """``likelihood(model(x))``: (mean[T, N], variance[T, N]) as read at Bayesian7.py:558-560 / :668-670."""
        xs = self.transform_inputs(x_unit).contiguous()
        out = self._per_task(lambda eng: eng.posterior(xs, min_variance))
        return torch.stack([o[0] for o in out]), torch.stack([o[1] for o in out])

    def _per_task(self, fn):
        """Run ``fn(engine)`` for every task, each on its own stream (forked from / joined to the current stream)."""
        if not self._streams:
            return [fn(eng) for eng in self.engines]
        cur = torch.cuda.current_stream(self.device)
        for st in self._streams:
            st.wait_stream(cur)

        def run(eng, st):
            torch.cuda.set_device(self.device)          # the current device is per host thread
            with torch.cuda.stream(st):
                return fn(eng)
        if len(self.engines) > 1:
            # one host thread per task: a sliced sweep reads its count of guard-flagged candidates back (one stream synchronise
            # per call), which would serialise the T task sweeps if one thread issued them; the C ABI calls release the GIL
            if self._pool is None:
                from concurrent.futures import ThreadPoolExecutor
                self._pool = ThreadPoolExecutor(max_workers=len(self.engines), thread_name_prefix="bo_svgp_task")
            out = list(self._pool.map(run, self.engines, self._streams))
        else:
            out = [run(self.engines[0], self._streams[0])]
        for st in self._streams:
            cur.wait_stream(st)
        for o in out:                                   # allocated on the side streams, consumed on the current one
            for t in (o if isinstance(o, (tuple, list)) else (o,)):
                if isinstance(t, torch.Tensor):
                    t.record_stream(cur)
        return out

    def variance_score(self, x_unit, min_variance: float = MIN_VARIANCE):
        """Uncertainty score of the pool scan: ``pred.variance.sum(dim=0)`` (Bayesian7.py:670-671), on the device."""
        xs = self.transform_inputs(x_unit).contiguous()
        vars_ = self._per_task(lambda eng: eng.sweep("var", candidates=xs, topk=0, min_variance=min_variance, return_all=True)[3])
        score = vars_[0].clone()
        for v in vars_[1:]:
            score.add_(v)
        return score

    def select_batch(self, cand_unit, batch_k: int, K_big_cap: int = 8000, fps_start: int = 0,
                     min_variance: float = MIN_VARIANCE):
        """top-K_big by summed variance -> farthest-point sampling of ``batch_k`` points (Bayesian7.py:673-688),
        without the score D2H, the CPU ``topk`` and the CPU FPS.  Returns (points[batch_k, d], pool indices)."""
        cand_unit = torch.as_tensor(cand_unit, dtype=torch.float64, device=self.device)
        N = cand_unit.shape[0]
        score = self.variance_score(cand_unit, min_variance)
        K_big = int(min(max(5000, 20 * batch_k), K_big_cap, N))
        _, idx_big = self.engines[0].topk_scores(score, K_big)                  # device radix select (bo_topk_scores)
        big = cand_unit[idx_big].contiguous()
        if batch_k >= K_big:
            return big, idx_big
        sel = self.engines[0].fps(big, int(batch_k), int(fps_start))
        return big[sel], idx_big[sel]

    def close(self):
        if self._pool is not None:
            self._pool.shutdown(wait=True)
            self._pool = None
        for e in self.engines:
            e.close()
        self.engines = []
