"""GPEngine -- thin PyTorch-tensor front end of the C ABI (include/bo_b200.h).

PyTorch is plumbing here (device memory, streams); all arithmetic happens in ``libbo_b200.so``.
Takes the place of the botorch/gpytorch objects the reference drives:
``SingleTaskGP`` + prediction caches (optimization/Bayesian.py:89-94), ``model.posterior``
(optimization/Bayesian2.py:168-171, Bayesian6.py:615-617), the pool scan + top-K of
optimization/Bayesian7.py:664-682 and ``optimize_acqf`` (optimization/Bayesian.py:105-112).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import (ACQ_EI, ACQ_LOGEI, ACQ_MEAN, ACQ_UCB, ACQ_VAR, BO_MAX_TOPK, KERNEL_LINEAR_MATERN52, KERNEL_MATERN52,
                   KERNEL_RBF, BoLibraryError, BoSobol)
from .sobol import sobol_state

MIN_VARIANCE = 1e-6   # gpytorch settings.min_variance (double)

_KERNELS = {"matern52": KERNEL_MATERN52, "matern": KERNEL_MATERN52, "rbf": KERNEL_RBF,
            "linear_matern52": KERNEL_LINEAR_MATERN52, "linear+matern52": KERNEL_LINEAR_MATERN52,
            KERNEL_MATERN52: KERNEL_MATERN52, KERNEL_RBF: KERNEL_RBF, KERNEL_LINEAR_MATERN52: KERNEL_LINEAR_MATERN52}
SWEEP_MODES = {"auto": 0, "fp64": 1, "i8x7": 2, "i8x8": 3}   # BO_SWEEP_* of include/bo_b200.h
_ACQS = {"ei": ACQ_EI, "logei": ACQ_LOGEI, "ucb": ACQ_UCB, "var": ACQ_VAR, "mean": ACQ_MEAN,
         ACQ_EI: ACQ_EI, ACQ_LOGEI: ACQ_LOGEI, ACQ_UCB: ACQ_UCB, ACQ_VAR: ACQ_VAR, ACQ_MEAN: ACQ_MEAN}


class NotPositiveDefiniteError(RuntimeError):
    """bo_fit / bo_append status k > 0: Cholesky broke down at (1-based) pivot k.  The caller retries with
    more jitter, as optimization/Bayesian6.py:482-488 does."""

    def __init__(self, pivot: int):
        super().__init__(f"kernel matrix not positive definite at pivot {pivot}")
        self.pivot = int(pivot)


class BoError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libbo_b200 error {code}: {msg}")
        self.code = code


def _stream_ptr(device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class GPEngine:
    """One fitted exact GP resident on one B200 plus its sweep workspaces (opaque ``bo_handle``)."""

    def __init__(self, device: Optional[torch.device] = None):
        self._lib = _lib.load()
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        if device is None or torch.device(device).type != "cuda":
            raise BoLibraryError("GPEngine needs a CUDA (B200, sm_100) device; there is no CPU fallback")
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        h = C.c_void_p()
        rc = self._lib.bo_create(C.byref(h), int(self.device.index))
        if rc != 0:
            raise BoError(rc, "bo_create failed (no usable sm_100 device?)")
        self._h = h
        self.sweep_mode = "auto"
        self.n = 0
        self.d = 0

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.bo_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc < 0:
            raise BoError(rc, self._lib.bo_last_error(self._h).decode())
        return rc

    def _linear_variance(self, linear_variance, d: int) -> float:
        """Scalar LinearKernel variance, or one per input dimension (gpytorch ``LinearKernel(ard_num_dims=d)``): the latter
        is handed to the library for the next fit / load (bo_set_linear_variance_ard) and its mean returned as the scalar."""
        v = torch.as_tensor(linear_variance, dtype=torch.float64).reshape(-1).cpu()
        if v.numel() == 1:
            self._check(self._lib.bo_set_linear_variance_ard(self._h, None, 0))
            return float(v[0])
        if v.numel() != d:
            raise ValueError(f"linear_variance must be a scalar or have d = {d} entries, got {v.numel()}")
        arr = (C.c_double * d)(*v.tolist())
        self._check(self._lib.bo_set_linear_variance_ard(self._h, arr, d))
        return float(v.mean())

    def _dev64(self, t, shape=None) -> torch.Tensor:
        t = torch.as_tensor(t, dtype=torch.float64, device=self.device).contiguous()
        if shape is not None:
            t = t.reshape(shape)
        return t

    # -- fit ----------------------------------------------------------------------------------
    def fit(self, X, y, kernel="matern52", lengthscale=1.0, outputscale=1.0, noise=1e-3, mean=0.0, jitter=0.0,
            linear_variance=0.0):
        """K = k(X,X) + (noise+jitter) I ; L ; alpha ; explicit L^-1 (optimization/Bayesian.py:89-94).
        ``kernel="linear_matern52"`` is ScaleKernel(Linear + Matern-5/2) of Bayesian6.py:471-473 / Bayesian7.py:162-166
        with LinearKernel variance ``linear_variance``.  CPU tensors go through the host-buffer entry."""
        X = torch.as_tensor(X)
        n, d = X.shape
        ls = torch.as_tensor(lengthscale, dtype=torch.float64).reshape(-1).cpu()
        if ls.numel() == 1:
            ls = ls.repeat(d)
        if ls.numel() != d:
            raise ValueError("lengthscale must be a scalar or have d entries")
        ls_arr = (C.c_double * d)(*ls.tolist())
        host = X.device.type == "cpu"
        if host:
            Xb = X.to(torch.float64).contiguous()
            yb = torch.as_tensor(y, dtype=torch.float64).reshape(-1).contiguous()
        else:
            Xb = self._dev64(X)
            yb = self._dev64(y, (-1,))
        if yb.numel() != n:
            raise ValueError("y must have n entries")
        linear_variance = self._linear_variance(linear_variance, d) if _KERNELS[kernel] == KERNEL_LINEAR_MATERN52 else 0.0
        with torch.cuda.device(self.device):
            rc = self._lib.bo_fit_ex(self._h, Xb.data_ptr(), yb.data_ptr(), n, d, _KERNELS[kernel], ls_arr, float(outputscale),
                                     float(noise), float(mean), float(jitter), float(linear_variance), 1 if host else 0,
                                     _stream_ptr(self.device))
        self._check(rc)
        if rc > 0:
            raise NotPositiveDefiniteError(rc)
        self.n, self.d = n, d
        return self

    def state(self):
        """(alpha[n], L[n,n], L^-1[n,n]) as device tensors (introspection for the parity tests)."""
        n = self.n
        alpha = torch.empty(n, dtype=torch.float64, device=self.device)
        L = torch.empty(n, n, dtype=torch.float64, device=self.device)
        Li = torch.empty(n, n, dtype=torch.float64, device=self.device)
        self._check(self._lib.bo_get_state(self._h, alpha.data_ptr(), L.data_ptr(), Li.data_ptr(), _stream_ptr(self.device)))
        return alpha, L, Li

    # -- posterior ------------------------------------------------------------------------------
    def posterior(self, Xs, min_variance: float = MIN_VARIANCE):
        """(mean[N], variance[N]) -- model.posterior(X).mean/.variance (Bayesian2.py:168-171)."""
        Xs = self._dev64(Xs).reshape(-1, self.d)
        N = Xs.shape[0]
        mean = torch.empty(N, dtype=torch.float64, device=self.device)
        var = torch.empty(N, dtype=torch.float64, device=self.device)
        self._check(self._lib.bo_posterior(self._h, Xs.data_ptr(), N, float(min_variance), mean.data_ptr(),
                                           var.data_ptr(), _stream_ptr(self.device)))
        return mean, var

    def load_svgp(self, Z, var_mean, var_chol, kernel="linear_matern52", lengthscale=1.0, outputscale=1.0,
                  linear_variance=0.0, mean=0.0, noise=0.0, jitter=1e-6):
        """Load one task of a sparse variational GP (whitened VariationalStrategy + CholeskyVariationalDistribution,
        optimization/Bayesian7.py:129-195): afterwards ``posterior`` / ``sweep`` evaluate its predictive distribution
        (mean = c + u^T m, var = k** + jitter - |u|^2 + |Ls^T u|^2 + noise with u = L^-1 k(Z, x*))."""
        Z = self._dev64(Z)
        M, d = Z.shape
        m = self._dev64(var_mean, (-1,))
        Ls = self._dev64(var_chol).reshape(M, M)
        if m.numel() != M:
            raise ValueError("var_mean must have M entries")
        ls = torch.as_tensor(lengthscale, dtype=torch.float64).reshape(-1).cpu()
        if ls.numel() == 1:
            ls = ls.repeat(d)
        if ls.numel() != d:
            raise ValueError("lengthscale must be a scalar or have d entries")
        ls_arr = (C.c_double * d)(*ls.tolist())
        linear_variance = self._linear_variance(linear_variance, d) if _KERNELS[kernel] == KERNEL_LINEAR_MATERN52 else 0.0
        with torch.cuda.device(self.device):
            rc = self._lib.bo_svgp_load(self._h, Z.data_ptr(), M, d, _KERNELS[kernel], ls_arr, float(outputscale),
                                        float(linear_variance), float(mean), float(noise), float(jitter), m.data_ptr(),
                                        Ls.data_ptr(), _stream_ptr(self.device))
        self._check(rc)
        if rc > 0:
            raise NotPositiveDefiniteError(rc)
        self.n, self.d = M, d
        return self

    def posterior_multi(self, Y, Xs, means=None, min_variance: float = MIN_VARIANCE, with_variance: bool = True):
        """(mean[N,m], variance[N]) of m outputs Y (n,m) sharing the fitted kernel matrix (one Cholesky, m right-hand sides)."""
        Y = self._dev64(Y).reshape(self.n, -1)
        m = Y.shape[1]
        Xs = self._dev64(Xs).reshape(-1, self.d)
        N = Xs.shape[0]
        mean = torch.empty(N, m, dtype=torch.float64, device=self.device)
        var = torch.empty(N, dtype=torch.float64, device=self.device) if with_variance else None
        mh = None
        if means is not None:
            mv = [float(v) for v in torch.as_tensor(means, dtype=torch.float64).reshape(-1).tolist()]
            mh = (C.c_double * m)(*mv)
        self._check(self._lib.bo_posterior_multi(self._h, Y.data_ptr(), m, mh, Xs.data_ptr(), N, float(min_variance),
                                                 mean.data_ptr(), _ptr(var), _stream_ptr(self.device)))
        return mean, var

    # -- sweep ----------------------------------------------------------------------------------
    def sweep(self, acq="ei", best_f=0.0, beta=2.0, candidates=None, sobol: Optional[BoSobol] = None,
              first_index: int = 0, count: Optional[int] = None, topk: int = 1,
              min_variance: float = MIN_VARIANCE, return_all: bool = False):
        """Score a pool shard and return (values[topk], global indices[topk]) on the device.

        ``candidates`` (N,d) is an explicit pool shard; otherwise ``sobol`` (see ``sobol_state``) generates
        candidates ``first_index .. first_index+count`` inside the kernel.  ``return_all`` adds the dense
        (mean, var, acq) arrays.  Pool scan + top-K of optimization/Bayesian7.py:664-682.
        """
        if not 0 <= topk <= BO_MAX_TOPK:
            raise ValueError(f"topk must be in [0, {BO_MAX_TOPK}]")
        cand = None
        if candidates is not None:
            cand = self._dev64(candidates).reshape(-1, self.d)
            N = cand.shape[0]
        else:
            if sobol is None or count is None:
                raise ValueError("either candidates or (sobol, count) must be given")
            N = int(count)
        vals = torch.empty(max(topk, 1), dtype=torch.float64, device=self.device)
        idx = torch.empty(max(topk, 1), dtype=torch.int64, device=self.device)
        mean = var = av = None
        if return_all:
            mean = torch.empty(N, dtype=torch.float64, device=self.device)
            var = torch.empty(N, dtype=torch.float64, device=self.device)
            av = torch.empty(N, dtype=torch.float64, device=self.device)
        self._check(self._lib.bo_sweep(
            self._h, _ACQS[acq], float(best_f), float(beta), float(min_variance), _ptr(cand),
            None if cand is not None else C.byref(sobol), int(first_index), N, int(topk),
            vals.data_ptr(), idx.data_ptr(), _ptr(mean), _ptr(var), _ptr(av), _stream_ptr(self.device)))
        vals, idx = vals[:topk], idx[:topk]
        if return_all:
            return vals, idx, mean, var, av
        return vals, idx

    def sweep_host(self, acq="ei", best_f=0.0, beta=2.0, candidates=None, sobol: Optional[BoSobol] = None,
                   first_index: int = 0, count: Optional[int] = None, topk: int = 1,
                   min_variance: float = MIN_VARIANCE):
        """Host-buffer entry (bo_sweep_host): host candidates in (or Sobol state), host top-k out; synchronous."""
        cand_ptr = None
        if candidates is not None:
            cand = torch.as_tensor(candidates, dtype=torch.float64, device="cpu").contiguous().reshape(-1, self.d)
            N, cand_ptr = cand.shape[0], cand.data_ptr()
        else:
            N = int(count)
        vals = torch.empty(topk, dtype=torch.float64)
        idx = torch.empty(topk, dtype=torch.int64)
        self._check(self._lib.bo_sweep_host(
            self._h, _ACQS[acq], float(best_f), float(beta), float(min_variance), cand_ptr,
            None if candidates is not None else C.byref(sobol), int(first_index), N, int(topk),
            vals.data_ptr(), idx.data_ptr(), _stream_ptr(self.device)))
        return vals, idx

    def sobol_points(self, sobol: BoSobol, idx) -> torch.Tensor:
        """Coordinates of pool points ``idx`` (Bayesian7.py:682 ``cand_unit_cpu[idxs_big]``)."""
        idx = torch.as_tensor(idx, dtype=torch.int64, device=self.device).contiguous().reshape(-1)
        out = torch.empty(idx.numel(), self.d, dtype=torch.float64, device=self.device)
        self._check(self._lib.bo_sobol_points(self._h, C.byref(sobol), idx.data_ptr(), idx.numel(), out.data_ptr(),
                                              _stream_ptr(self.device)))
        return out

    # -- refinement / append / LML ----------------------------------------------------------------
    def acq_grad(self, Xq, acq="ei", best_f=0.0, beta=2.0, min_variance: float = MIN_VARIANCE):
        Xq = self._dev64(Xq).reshape(-1, self.d)
        k = Xq.shape[0]
        val = torch.empty(k, dtype=torch.float64, device=self.device)
        grad = torch.empty(k, self.d, dtype=torch.float64, device=self.device)
        self._check(self._lib.bo_acq_grad(self._h, _ACQS[acq], float(best_f), float(beta), float(min_variance),
                                          Xq.data_ptr(), k, val.data_ptr(), grad.data_ptr(), _stream_ptr(self.device)))
        return val, grad

    def refine(self, starts, acq="ei", best_f=0.0, beta=2.0, iters: int = 50, min_variance: float = MIN_VARIANCE):
        """Batched projected-gradient refinement of the top-k starts (optimize_acqf stand-in, Bayesian.py:105-112)."""
        starts = self._dev64(starts).reshape(-1, self.d)
        k = starts.shape[0]
        x = torch.empty_like(starts)
        val = torch.empty(k, dtype=torch.float64, device=self.device)
        self._check(self._lib.bo_refine(self._h, _ACQS[acq], float(best_f), float(beta), float(min_variance),
                                        starts.data_ptr(), k, int(iters), x.data_ptr(), val.data_ptr(),
                                        _stream_ptr(self.device)))
        return x, val

    def append(self, x, y: Optional[float] = None):
        """Row-append update; ``y=None`` appends the Kriging-believer value mu(x)."""
        x = self._dev64(x).reshape(-1)
        rc = self._check(self._lib.bo_append(self._h, x.data_ptr(), 0.0 if y is None else float(y),
                                             1 if y is None else 0, _stream_ptr(self.device)))
        if rc > 0:
            raise NotPositiveDefiniteError(rc)
        self.n += 1
        return self

    def lml_grad_batched(self, X, y, thetas, kernel="matern52", mean: float = 0.0):
        """Batched exact LML and gradient w.r.t. log(lengthscale[d]), log(outputscale), log(noise) [, log(linear variance)]."""
        X = self._dev64(X)
        n, d = X.shape
        y = self._dev64(y, (-1,))
        p = d + 2 + (1 if _KERNELS[kernel] == KERNEL_LINEAR_MATERN52 else 0)      # + log linear variance
        th = torch.as_tensor(thetas, dtype=torch.float64, device="cpu").contiguous().reshape(-1, p)
        R = th.shape[0]
        lml = torch.empty(R, dtype=torch.float64)
        grad = torch.empty(R, p, dtype=torch.float64)
        status = torch.empty(R, dtype=torch.int32)
        self._check(self._lib.bo_lml_grad_batched(self._h, X.data_ptr(), y.data_ptr(), n, d, _KERNELS[kernel],
                                                  float(mean), th.data_ptr(), R, lml.data_ptr(), grad.data_ptr(),
                                                  status.data_ptr(), _stream_ptr(self.device)))
        return lml, grad, status

    def topk_scores(self, scores, k: int, first_index: int = 0):
        """(values[k], indices[k]) of the k <= 8192 best scores on the device, value desc / index asc / NaN last
        (the CPU ``torch.topk(unc, K_big)`` of Bayesian7.py:681)."""
        scores = self._dev64(scores, (-1,))
        vals = torch.empty(int(k), dtype=torch.float64, device=self.device)
        idx = torch.empty(int(k), dtype=torch.int64, device=self.device)
        self._check(self._lib.bo_topk_scores(self._h, scores.data_ptr(), scores.numel(), int(first_index), int(k),
                                             vals.data_ptr(), idx.data_ptr(), _stream_ptr(self.device)))
        return vals, idx

    def fps(self, X, m: int, start: int = 0) -> torch.Tensor:
        """Indices of m farthest-point samples of X (N,d), greedy from ``start`` (Bayesian7.py:82-107)."""
        X = self._dev64(X)
        N, d = X.shape
        idx = torch.empty(int(m), dtype=torch.int64, device=self.device)
        self._check(self._lib.bo_fps(self._h, X.data_ptr(), N, d, int(m), int(start), idx.data_ptr(), _stream_ptr(self.device)))
        return idx

    # -- diagnostics ------------------------------------------------------------------------------
    def fp64_peak_tflops(self, use_dmma: bool = True, seconds: float = 0.3) -> float:
        out = C.c_double(0.0)
        self._check(self._lib.bo_fp64_peak(self._h, 1 if use_dmma else 0, float(seconds), C.byref(out)))
        return out.value

    def gemm_probe_tflops(self, m: int, n: int, k: int, cfg: int = 1, reps: int = 10) -> float:
        out = C.c_double(0.0)
        self._check(self._lib.bo_gemm_probe(self._h, m, n, k, cfg, reps, C.byref(out)))
        return out.value

    def i8_peak_tops(self, seconds: float = 0.3) -> float:
        out = C.c_double(0.0)
        self._check(self._lib.bo_i8_peak(self._h, float(seconds), C.byref(out)))
        return out.value

    def set_sweep_mode(self, mode: str = "auto"):
        """Contraction of the sweep's variance term: "auto" (INT8-sliced tensor path for large pools, FP64 DMMA otherwise),
        "fp64", "i8x7", "i8x8" (bo_set_sweep_mode)."""
        self._check(self._lib.bo_set_sweep_mode(self._h, SWEEP_MODES[mode]))
        self.sweep_mode = mode

    def resolve_sweep_mode(self, pool_total: int) -> str:
        """The pinned mode the current mode resolves to for a pool of ``pool_total`` candidates (bo_resolve_sweep_mode):
        sharded callers resolve once on the global pool size and pin the result for every shard."""
        rc = int(self._lib.bo_resolve_sweep_mode(self._h, int(pool_total)))
        self._check(min(rc, 0))
        return {v: k for k, v in SWEEP_MODES.items()}[rc]

    def last_sweep_path(self) -> int:
        """0 = FP64 DMMA, 7 / 8 = INT8-sliced with that many slices, -1 = no sweep yet."""
        return int(self._lib.bo_last_sweep_path(self._h))

    def last_sweep_flagged(self) -> int:
        """Candidates of the last sliced sweep its accuracy guard re-scored on the FP64 contraction."""
        return int(self._lib.bo_last_sweep_flagged(self._h))

    def launch_count(self) -> int:
        return int(self._lib.bo_launch_count(self._h))

    def last_sweep_ms(self) -> float:
        return float(self._lib.bo_last_sweep_ms(self._h))

    def release_workspace(self):
        self._check(self._lib.bo_release_workspace(self._h))
