"""FP64 CPU oracle for the GP surrogate + acquisition hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module, and only as the checker (or the timed CPU baseline) -- the product
(`bayesianoptimizer_b200`) never imports it and has no CPU fallback.

PARITY UNPINNED: the reference (billbearhunter/BayesianOptimizer) has no tests, golden vectors or
known-answer fixtures for this path, and its arithmetic lives in un-vendored, un-pinned third-party
packages (botorch / gpytorch / linear_operator -- versions unpinned: no requirements or lock file
exists; inferred botorch >= 0.12 from optimization/Bayesian.py:156) that are not installed here and
cannot be fetched.  This file restates their published exact-GP algorithm (exact Cholesky, i.e.
gpytorch under max_cholesky_size = inf) in NumPy/SciPy and is anchored on the reference's call sites:

* model construction            optimization/Bayesian.py:89-94  (SingleTaskGP + ExactMarginalLogLikelihood)
* explicit Matern-5/2 ARD kernel optimization/Bayesian6.py:470-473, optimization/Bayesian7.py:162-166
* input normalisation           optimization/Bayesian.py:137,164 ; optimization/Bayesian7.py:280
* outcome standardisation       optimization/Bayesian6.py:427-443 (botorch ``Standardize``: ddof=1)
* incumbent best_f              optimization/Bayesian.py:98 ; optimization/Bayesian2.py:221-227
* acquisition (EI semantics)    optimization/Bayesian.py:98-113
* pool sweep / top-K shape      optimization/Bayesian7.py:650-688
* Cholesky-failure -> jitter    optimization/Bayesian6.py:482-488

It is independently cross-checked in tests/ against scikit-learn's GaussianProcessRegressor and
scipy.stats.norm (separately written implementations that are present in the container).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import scipy.linalg as sla
import scipy.special as ssp

KERNEL_MATERN52 = 0
KERNEL_RBF = 1
KERNEL_LINEAR_MATERN52 = 2   # ScaleKernel(LinearKernel + MaternKernel(2.5, ard)): Bayesian6.py:471-473, Bayesian7.py:162-166

ACQ_EI = 0
ACQ_LOGEI = 1
ACQ_UCB = 2
ACQ_VAR = 3      # posterior-variance score (active-learning sweep, Bayesian7.py:670-671)
ACQ_MEAN = 4

MIN_VARIANCE = 1e-6   # gpytorch settings.min_variance (double) clamp applied by MultivariateNormal.variance
SQRT5 = math.sqrt(5.0)


# --------------------------------------------------------------------------------------
# transforms (botorch.utils.transforms.normalize / unnormalize, Standardize) -- Bayesian.py:137,164
# --------------------------------------------------------------------------------------
def normalize(X, bounds):
    """(X - lo) / (hi - lo); bounds is (2, d) like Bayesian.py:42."""
    bounds = np.asarray(bounds, dtype=np.float64)
    return (np.asarray(X, dtype=np.float64) - bounds[0]) / (bounds[1] - bounds[0])


def unnormalize(X, bounds):
    bounds = np.asarray(bounds, dtype=np.float64)
    return np.asarray(X, dtype=np.float64) * (bounds[1] - bounds[0]) + bounds[0]


def standardize(y):
    """botorch ``Standardize``: (y - mean) / std with unbiased std (ddof=1), std floored at 1e-8."""
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    mu = float(y.mean())
    sd = float(y.std(ddof=1)) if y.size > 1 else 1.0
    if not sd >= 1e-8:
        sd = 1.0
    return (y - mu) / sd, mu, sd


# --------------------------------------------------------------------------------------
# kernels (gpytorch MaternKernel(nu=2.5, ard) / RBFKernel(ard) under ScaleKernel)
# --------------------------------------------------------------------------------------
def scaled_sqdist(A, B, lengthscale):
    """Direct-difference ARD squared distance  sum_k ((a_k - b_k)/l_k)^2  (m, n)."""
    ls = np.asarray(lengthscale, dtype=np.float64).reshape(1, -1)
    As = np.asarray(A, dtype=np.float64) / ls
    Bs = np.asarray(B, dtype=np.float64) / ls
    out = np.zeros((As.shape[0], Bs.shape[0]))
    for k in range(As.shape[1]):
        diff = As[:, k:k + 1] - Bs[:, k].reshape(1, -1)
        out += diff * diff
    return out


def kernel_from_sqdist(sq, kind, outputscale):
    if kind == KERNEL_MATERN52:
        r = np.sqrt(sq)
        return outputscale * (1.0 + SQRT5 * r + (5.0 / 3.0) * sq) * np.exp(-SQRT5 * r)
    if kind == KERNEL_RBF:
        return outputscale * np.exp(-0.5 * sq)
    raise ValueError(f"unknown kernel kind {kind}")


def kernel_matrix(A, B, kind, lengthscale, outputscale, linear_variance=0.0):
    """k(A, B).  Kind 2 is ScaleKernel(LinearKernel + Matern-5/2): s2 * (v <a, b> + matern(a, b)); gpytorch's LinearKernel
    acts on the raw inputs; its variance is a scalar, or one per input dimension with ard_num_dims (Bayesian7.py:162-166 asks
    for that: raw_variance (T, 1, d)), in which case v <a, b> reads sum_k v_k a_k b_k."""
    if kind == KERNEL_LINEAR_MATERN52:
        A = np.asarray(A, dtype=np.float64); B = np.asarray(B, dtype=np.float64)
        v = np.broadcast_to(np.asarray(linear_variance, dtype=np.float64), (A.shape[1],))     # scalar, or one per dimension (ARD)
        lin = np.zeros((A.shape[0], B.shape[0]))
        for k in range(A.shape[1]):
            lin += v[k] * A[:, k:k + 1] * B[:, k].reshape(1, -1)
        return outputscale * lin + kernel_from_sqdist(scaled_sqdist(A, B, lengthscale), KERNEL_MATERN52, outputscale)
    return kernel_from_sqdist(scaled_sqdist(A, B, lengthscale), kind, outputscale)


def prior_variance(X, kind, outputscale, linear_variance=0.0):
    """k(x, x) per row: s2 for the stationary kinds, s2 (v |x|^2 + 1) for linear + Matern."""
    X = np.asarray(X, dtype=np.float64)
    if kind == KERNEL_LINEAR_MATERN52:
        v = np.broadcast_to(np.asarray(linear_variance, dtype=np.float64), (X.shape[1],))
        return outputscale * (np.sum(v[None, :] * X * X, axis=1) + 1.0)
    return np.full(X.shape[0], float(outputscale))


class NotPositiveDefinite(np.linalg.LinAlgError):
    """Cholesky failed at (1-based) pivot ``pivot`` -- mirrors the retry-with-jitter convention
    of Bayesian6.py:482-488."""

    def __init__(self, pivot):
        super().__init__(f"matrix not positive definite at pivot {pivot}")
        self.pivot = int(pivot)


@dataclass
class GPFit:
    X: np.ndarray
    y: np.ndarray
    kind: int
    lengthscale: np.ndarray
    outputscale: float
    noise: float
    mean: float
    L: np.ndarray          # lower Cholesky factor of K + (noise + jitter) I
    alpha: np.ndarray      # (K + noise I)^-1 (y - mean)
    linear_variance: float = 0.0

    @property
    def n(self):
        return self.X.shape[0]

    @property
    def d(self):
        return self.X.shape[1]


def _cholesky_lower(K):
    L, info = sla.lapack.dpotrf(K, lower=1, clean=1, overwrite_a=0)
    if info > 0:
        raise NotPositiveDefinite(info)
    if info < 0:
        raise ValueError(f"dpotrf illegal argument {-info}")
    return L


def fit(X, y, kind=KERNEL_MATERN52, lengthscale=None, outputscale=1.0, noise=1e-3, mean=0.0, jitter=0.0,
        linear_variance=0.0):
    """K = k(X,X) + (noise + jitter) I ; L = chol(K) ; alpha = K^-1 (y - mean).  """
    X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    n, d = X.shape
    ls = np.full(d, 1.0) if lengthscale is None else np.broadcast_to(np.asarray(lengthscale, dtype=np.float64), (d,)).copy()
    K = kernel_matrix(X, X, kind, ls, outputscale, linear_variance)
    K[np.diag_indices(n)] = prior_variance(X, kind, outputscale, linear_variance) + noise + jitter   # exact diagonal (d(x,x)=0)
    L = _cholesky_lower(K)
    alpha = sla.cho_solve((L, True), y - mean)
    lv = float(linear_variance) if np.ndim(linear_variance) == 0 else np.asarray(linear_variance, dtype=np.float64).copy()
    return GPFit(X, y, kind, ls, float(outputscale), float(noise), float(mean), L, alpha, lv)


def posterior(gp: GPFit, Xs, min_variance=MIN_VARIANCE, chunk=2048):
    """mean = m + k*^T alpha ; var = max(s2 - ||L^-1 k*||^2, min_variance).  Chunked like Bayesian7.py:63,665."""
    Xs = np.ascontiguousarray(Xs, dtype=np.float64).reshape(-1, gp.d)
    N = Xs.shape[0]
    mu = np.empty(N)
    var = np.empty(N)
    for s in range(0, N, chunk):
        Ks = kernel_matrix(gp.X, Xs[s:s + chunk], gp.kind, gp.lengthscale, gp.outputscale, gp.linear_variance)   # (n, c)
        mu[s:s + chunk] = gp.mean + Ks.T @ gp.alpha
        V = sla.solve_triangular(gp.L, Ks, lower=True, check_finite=False)
        var[s:s + chunk] = prior_variance(Xs[s:s + chunk], gp.kind, gp.outputscale, gp.linear_variance) - np.einsum("ij,ij->j", V, V)
    return mu, np.maximum(var, min_variance)


# --------------------------------------------------------------------------------------
# analytic acquisition (botorch.acquisition.analytic semantics)
# --------------------------------------------------------------------------------------
_LOG_SQRT_2PI = 0.5 * math.log(2.0 * math.pi)
_HALF_LOG_PI_2 = 0.5 * math.log(math.pi / 2.0)
_INV_SQRT2 = 1.0 / math.sqrt(2.0)


def _phi(u):
    return np.exp(-0.5 * u * u) / math.sqrt(2.0 * math.pi)


def _Phi(u):
    return 0.5 * ssp.erfc(-u * _INV_SQRT2)


def _log1mexp(x):
    """log(1 - exp(x)) for x < 0, stable (Maechler 2012)."""
    x = np.asarray(x, dtype=np.float64)
    return np.where(x > -math.log(2.0), np.log(-np.expm1(x)), np.log1p(-np.exp(x)))


def log_ei_helper(u):
    """log(phi(u) + u Phi(u)), accurate for u -> -inf (erfcx form, botorch `_log_ei_helper`)."""
    u = np.asarray(u, dtype=np.float64)
    out = np.empty_like(u)
    hi = u > -1.0
    uh = u[hi]
    out[hi] = np.log(_phi(uh) + uh * _Phi(uh))
    ul = u[~hi]
    # phi(u) * (1 - |u| sqrt(pi/2) erfcx(|u|/sqrt2))
    w = np.log(ssp.erfcx(-ul * _INV_SQRT2) * np.abs(ul)) + _HALF_LOG_PI_2
    out[~hi] = -0.5 * ul * ul - _LOG_SQRT_2PI + _log1mexp(w)
    return out


def acquisition(mu, var, kind, best_f=0.0, beta=2.0):
    """Analytic EI / LogEI / UCB(beta) / variance / mean for a maximisation problem (Bayesian.py:98)."""
    mu = np.asarray(mu, dtype=np.float64)
    var = np.asarray(var, dtype=np.float64)
    if kind == ACQ_VAR:
        return var.copy()
    if kind == ACQ_MEAN:
        return mu.copy()
    sigma = np.sqrt(var)
    if kind == ACQ_UCB:
        return mu + math.sqrt(beta) * sigma
    u = (mu - best_f) / sigma
    if kind == ACQ_EI:
        return sigma * (_phi(u) + u * _Phi(u))
    if kind == ACQ_LOGEI:
        return np.log(sigma) + log_ei_helper(u)
    raise ValueError(f"unknown acquisition kind {kind}")


def topk(values, k, first_index=0):
    """Top-k by (value desc, index asc); NaN counts as -inf.  Returns (values, global int64 indices)."""
    v = np.asarray(values, dtype=np.float64)
    key = np.where(np.isnan(v), -np.inf, v)
    order = np.lexsort((np.arange(v.size), -key))[:k]
    return key[order], order.astype(np.int64) + int(first_index)


def merge_topk(vals_list, idx_list, k):
    """Merge per-shard top-k lists with the same (value desc, index asc) order -- the C1 reduce."""
    v = np.concatenate([np.asarray(a, dtype=np.float64) for a in vals_list])
    i = np.concatenate([np.asarray(a, dtype=np.int64) for a in idx_list])
    keep = i >= 0
    v, i = v[keep], i[keep]
    order = np.lexsort((i, -v))[:k]
    return v[order], i[order]


# --------------------------------------------------------------------------------------
# Sobol candidates: position-independent restatement of torch.quasirandom.SobolEngine.draw
# --------------------------------------------------------------------------------------
SOBOL_BITS = 30


def sobol_points(sobolstate, shift, first_index, count):
    """Point i = shift XOR (XOR over set bits b of gray(i) of V[:, b]) scaled by 2^-30.

    ``sobolstate`` (d, 30) and ``shift`` (d,) are the integer state of a fresh
    ``torch.quasirandom.SobolEngine(d, scramble=True, seed=s)``; point i equals row i of its
    ``draw`` output (dtype float64; torch keeps point 0 in float32, reproduced here).  Candidate pool generator for the sweep (Bayesian7.py:650-655
    uses an LHS pool; Bayesian.py:105-112 `optimize_acqf` raw samples are scrambled Sobol).
    """
    V = np.asarray(sobolstate, dtype=np.int64)
    sh = np.asarray(shift, dtype=np.int64)
    idx = np.arange(first_index, first_index + count, dtype=np.int64)
    gray = idx ^ (idx >> 1)
    acc = np.broadcast_to(sh, (count, V.shape[0])).copy()
    for b in range(SOBOL_BITS):
        bit = ((gray >> b) & 1).astype(bool)
        acc[bit] ^= V[:, b]
    pts = acc.astype(np.float64) * (2.0 ** -SOBOL_BITS)
    if first_index == 0 and count > 0:
        # torch stores point 0 (`_first_point`) in float32 before the cast to the draw dtype
        pts[0] = acc[0].astype(np.float32).astype(np.float64) * (2.0 ** -SOBOL_BITS)
    return pts


def sweep(gp: GPFit, Xs, acq_kind, best_f=0.0, beta=2.0, k=1, first_index=0, min_variance=MIN_VARIANCE):
    """Score a candidate pool and return (topk values, topk indices, mu, var, acq) -- Bayesian7.py:664-682 shape."""
    mu, var = posterior(gp, Xs, min_variance)
    a = acquisition(mu, var, acq_kind, best_f, beta)
    tv, ti = topk(a, k, first_index)
    return tv, ti, mu, var, a


# --------------------------------------------------------------------------------------
# exact marginal log likelihood + gradient (Bayesian.py:92-93)
# --------------------------------------------------------------------------------------
def lml_and_grad(X, y, kind, lengthscale, outputscale, noise, mean=0.0, linear_variance=0.0):
    """log N(y; m, K + noise I) and d/d(log l_k), d/d(log s2), d/d(log noise)  (un-normalised, no priors).

    Returns (lml, grad[d + 2]); for the linear + Matern kind grad has a last entry d/d(log v).
    Raises NotPositiveDefinite like ``fit``.
    """
    X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    n, d = X.shape
    ls = np.broadcast_to(np.asarray(lengthscale, dtype=np.float64), (d,)).copy()
    sq = scaled_sqdist(X, X, ls)
    np.fill_diagonal(sq, 0.0)
    lin_kind = kind == KERNEL_LINEAR_MATERN52
    Kf = kernel_from_sqdist(sq, KERNEL_MATERN52 if lin_kind else kind, outputscale)
    if lin_kind:
        Klin = outputscale * linear_variance * (X @ X.T)
        Kf = Kf + Klin
    K = Kf.copy()
    K[np.diag_indices(n)] = np.diag(Kf) + noise
    L = _cholesky_lower(K)
    r = y - mean
    alpha = sla.cho_solve((L, True), r)
    lml = -0.5 * float(r @ alpha) - float(np.log(np.diag(L)).sum()) - 0.5 * n * math.log(2.0 * math.pi)
    Kinv = sla.cho_solve((L, True), np.eye(n))
    W = np.outer(alpha, alpha) - Kinv            # dLML/dtheta = 0.5 tr(W dK/dtheta)
    grad = np.empty(d + 3 if lin_kind else d + 2)
    if kind == KERNEL_MATERN52 or lin_kind:
        rr = np.sqrt(sq)
        G = outputscale * (5.0 / 3.0) * (1.0 + SQRT5 * rr) * np.exp(-SQRT5 * rr)   # dk/d(sq) * -2 ... see below
    else:
        G = Kf                                                                     # rbf: dk/dl_k = k * D_k^2 / l_k^3
    WG = W * G
    for k in range(d):
        diff = (X[:, k:k + 1] - X[:, k].reshape(1, -1)) / ls[k]
        # dk/d(log l_k) = G * (D_k/l_k)^2   (Matern-5/2: s2 5/3 (1+sqrt5 r) e^{-sqrt5 r} D^2/l^2 ; RBF: k D^2/l^2)
        grad[k] = 0.5 * float(np.sum(WG * diff * diff))
    grad[d] = 0.5 * float(np.sum(W * Kf))                       # d/d(log s2): dK/d(log s2) = Kf
    grad[d + 1] = 0.5 * noise * float(np.trace(W))              # d/d(log noise): dK/d(log noise) = noise I
    if lin_kind:
        grad[d + 2] = 0.5 * float(np.sum(W * Klin))             # d/d(log v): dK/d(log v) = s2 v X X^T
    return lml, grad


# --------------------------------------------------------------------------------------
# Kriging-believer row append
# --------------------------------------------------------------------------------------
def append_point(gp: GPFit, x, y=None):
    """Border L with the new row; y=None appends the believer value mu(x) (then alpha' = [alpha; 0])."""
    x = np.asarray(x, dtype=np.float64).reshape(1, -1)
    k = kernel_matrix(gp.X, x, gp.kind, gp.lengthscale, gp.outputscale, gp.linear_variance)[:, 0]
    l = sla.solve_triangular(gp.L, k, lower=True, check_finite=False)
    lam2 = float(prior_variance(x, gp.kind, gp.outputscale, gp.linear_variance)[0]) + gp.noise - float(l @ l)
    if not lam2 > 0.0:
        raise NotPositiveDefinite(gp.n + 1)
    lam = math.sqrt(lam2)
    n = gp.n
    L2 = np.zeros((n + 1, n + 1))
    L2[:n, :n] = gp.L
    L2[n, :n] = l
    L2[n, n] = lam
    if y is None:
        y = gp.mean + float(k @ gp.alpha)
    X2 = np.vstack([gp.X, x])
    y2 = np.concatenate([gp.y, [y]])
    alpha2 = sla.cho_solve((L2, True), y2 - gp.mean)
    return GPFit(X2, y2, gp.kind, gp.lengthscale, gp.outputscale, gp.noise, gp.mean, L2, alpha2, gp.linear_variance)


# --------------------------------------------------------------------------------------
# acquisition gradient + projected refinement of starts (optimize_acqf stand-in, Bayesian.py:105-112)
# --------------------------------------------------------------------------------------
def posterior_with_grad(gp: GPFit, x):
    """mu, var (unclamped), dmu/dx, dvar/dx at one point x (d,)."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    ls = gp.lengthscale
    diff = (x[None, :] - gp.X) / ls[None, :]                 # (n, d)  (x - X_j)/l
    sq = np.sum(diff * diff, axis=1)
    if gp.kind in (KERNEL_MATERN52, KERNEL_LINEAR_MATERN52):
        r = np.sqrt(sq)
        e = np.exp(-SQRT5 * r)
        k = gp.outputscale * (1.0 + SQRT5 * r + (5.0 / 3.0) * sq) * e
        g = -gp.outputscale * (5.0 / 3.0) * (1.0 + SQRT5 * r) * e          # dk/d(sq) * 2
    else:
        k = gp.outputscale * np.exp(-0.5 * sq)
        g = -k
    dk = g[:, None] * diff / ls[None, :]                     # (n, d) dk_j/dx
    prior, dprior = gp.outputscale, 0.0
    if gp.kind == KERNEL_LINEAR_MATERN52:
        # ScaleKernel(Linear + Matern) of optimization/Bayesian6.py:470-478: k = s2 (v <x, x'> + matern)
        v = np.broadcast_to(np.asarray(gp.linear_variance, dtype=np.float64), x.shape)      # scalar or per dimension
        k = k + gp.outputscale * (gp.X @ (v * x))
        dk = dk + gp.outputscale * v[None, :] * gp.X
        prior = gp.outputscale * (float(np.sum(v * x * x)) + 1.0)
        dprior = 2.0 * gp.outputscale * v * x
    mu = gp.mean + float(k @ gp.alpha)
    w = sla.cho_solve((gp.L, True), k)                       # K^-1 k
    var = prior - float(k @ w)
    dmu = dk.T @ gp.alpha
    dvar = dprior - 2.0 * (dk.T @ w)
    return mu, var, dmu, dvar


def acquisition_with_grad(gp: GPFit, x, kind, best_f=0.0, beta=2.0, min_variance=MIN_VARIANCE):
    mu, var, dmu, dvar = posterior_with_grad(gp, x)
    if var < min_variance:
        var, dvar = min_variance, np.zeros_like(dvar)
    sigma = math.sqrt(var)
    dsig = dvar / (2.0 * sigma)
    if kind == ACQ_UCB:
        return mu + math.sqrt(beta) * sigma, dmu + math.sqrt(beta) * dsig
    if kind == ACQ_VAR:
        return var, dvar
    if kind == ACQ_MEAN:
        return mu, dmu
    u = (mu - best_f) / sigma
    ph, Ph = float(_phi(np.float64(u))), float(_Phi(np.float64(u)))
    if kind == ACQ_EI:
        # dEI = Phi(u) dmu + phi(u) dsigma
        return sigma * (ph + u * Ph), Ph * dmu + ph * dsig
    if kind == ACQ_LOGEI:
        h = float(log_ei_helper(np.array([u]))[0])
        # d log h/du = Phi/(phi + u Phi) computed stably as exp(log Phi - log h)
        logPhi = float(ssp.log_ndtr(u))
        dlogh = math.exp(logPhi - h)
        du = (dmu - u * dsig) / sigma
        return math.log(sigma) + h, dsig / sigma + dlogh * du
    raise ValueError(kind)


# --------------------------------------------------------------------------------------
# farthest-point sampling (optimization/Bayesian7.py:82-107) with a fixed start index
# --------------------------------------------------------------------------------------
def fps(X, m, start=0):
    """Greedy FPS: arg-max of the running minimum squared distance, first index on ties (torch.argmax)."""
    X = np.asarray(X, dtype=np.float64)
    idx = [int(start)]
    dist = np.full(X.shape[0], np.inf)
    for _ in range(1, m):
        diff = X - X[idx[-1]]
        s = np.zeros(X.shape[0])
        for k in range(X.shape[1]):
            s += diff[:, k] * diff[:, k]
        dist = np.minimum(dist, s)
        idx.append(int(np.argmax(dist)))
    return np.array(idx, dtype=np.int64)


# --------------------------------------------------------------------------------------
# several outputs sharing one kernel matrix (one Cholesky, m right-hand sides): SURVEY 8f N4
# --------------------------------------------------------------------------------------
def posterior_multi(gp: GPFit, Y, Xs, means=None, min_variance=MIN_VARIANCE):
    """Posterior means (N, m) of m independent outputs Y (n, m) that share gp's kernel matrix, plus the shared
    variance (N,).  The exact-GP analogue of the batched 8-task models (Bayesian1.py:109-113, Bayesian7.py:129-195)."""
    Y = np.asarray(Y, dtype=np.float64).reshape(gp.n, -1)
    m = Y.shape[1]
    means = np.zeros(m) if means is None else np.asarray(means, dtype=np.float64)
    A = sla.cho_solve((gp.L, True), Y - means[None, :])
    Xs = np.ascontiguousarray(Xs, dtype=np.float64).reshape(-1, gp.d)
    Ks = kernel_matrix(gp.X, Xs, gp.kind, gp.lengthscale, gp.outputscale, gp.linear_variance)
    mu = Ks.T @ A + means[None, :]
    _, var = posterior(gp, Xs, min_variance)
    return mu, var


# --------------------------------------------------------------------------------------
# log + standardise output transform and its lognormal back-transform (Bayesian6.py:427-443, 631-633, 703-707)
# --------------------------------------------------------------------------------------
@dataclass
class LogStandardize:
    shift: float
    mean: np.ndarray      # (1, m) mean of log(Y + shift)
    std: np.ndarray       # (1, m) unbiased std of log(Y + shift), floored at 1e-12

    @staticmethod
    def fit(Y):
        Y = np.asarray(Y, dtype=np.float64).reshape(len(Y), -1)
        eps = max(1e-12, float(np.abs(Y).max()) * 1e-6) if Y.size else 1e-6        # _compute_safe_epsilon, :421-425
        ymin = float(Y.min())
        shift = (-ymin + eps) if ymin <= 0.0 else eps                              # :431-435
        lg = np.log(Y + shift)
        std = lg.std(axis=0, ddof=1, keepdims=True)
        std = np.where(std < 1e-12, 1e-12, std)                                    # :440-442
        return LogStandardize(shift, lg.mean(axis=0, keepdims=True), std)

    def forward(self, Y):
        Y = np.asarray(Y, dtype=np.float64).reshape(len(Y), -1)
        return np.nan_to_num((np.log(Y + self.shift) - self.mean) / self.std, nan=0.0)   # :464-468

    def inverse_mean(self, mean_std, var_std):
        """E[Y] of the lognormal: exp(mu_log + var_log / 2) - shift (:631-633)."""
        log_mean = np.asarray(mean_std) * self.std + self.mean
        log_var = np.asarray(var_std) * self.std ** 2
        return np.exp(log_mean + 0.5 * log_var) - self.shift


# --------------------------------------------------------------------------------------
# SVGP predictive distribution: one task of the batched sparse variational GP of
# optimization/Bayesian7.py:129-195 evaluated as the pool scan of :664-671 does.  gpytorch semantics [3P-recall]:
# VariationalStrategy (whitened) -- L = chol(K_uu + jitter I), interp = L^-1 K_u*, mean = interp^T m + c,
# covar = (K_** + jitter I) + interp^T (S - I) interp, S = Ls Ls^T from CholeskyVariationalDistribution (lower triangle of
# chol_variational_covar); GaussianLikelihood adds the task's noise; MultivariateNormal.variance clamps at min_variance.
# --------------------------------------------------------------------------------------
@dataclass
class SVGPTask:
    Z: np.ndarray              # (M, d) inducing points, in the model's input space
    kind: int
    lengthscale: np.ndarray    # (d,)
    outputscale: float
    linear_variance: float
    mean: float                # ConstantMean
    noise: float               # likelihood noise (0: latent variance)
    jitter: float              # variational_cholesky_jitter: 1e-6 double / 1e-4 float
    m: np.ndarray              # (M,) whitened variational mean
    Ls: np.ndarray             # (M, M) chol_variational_covar (lower triangle used)


def svgp_predict(task: SVGPTask, Xs, min_variance=MIN_VARIANCE):
    Z = np.ascontiguousarray(task.Z, dtype=np.float64)
    M, d = Z.shape
    ls = np.broadcast_to(np.asarray(task.lengthscale, dtype=np.float64), (d,))
    Xs = np.ascontiguousarray(Xs, dtype=np.float64).reshape(-1, d)
    Kuu = kernel_matrix(Z, Z, task.kind, ls, task.outputscale, task.linear_variance)
    Kuu[np.diag_indices(M)] = prior_variance(Z, task.kind, task.outputscale, task.linear_variance) + task.jitter
    L = _cholesky_lower(Kuu)
    Kus = kernel_matrix(Z, Xs, task.kind, ls, task.outputscale, task.linear_variance)          # (M, N)
    interp = sla.solve_triangular(L, Kus, lower=True, check_finite=False)
    Ls = np.tril(np.asarray(task.Ls, dtype=np.float64))
    mean = task.mean + interp.T @ np.asarray(task.m, dtype=np.float64)
    w = Ls.T @ interp
    var = (prior_variance(Xs, task.kind, task.outputscale, task.linear_variance) + task.jitter
           - np.einsum("ij,ij->j", interp, interp) + np.einsum("ij,ij->j", w, w) + task.noise)
    return mean, np.maximum(var, min_variance)


def svgp_transform_inputs(x_unit, bounds, x_log_mean, x_log_std):
    """unit cube -> physical -> log -> standardised (BatchSVGP._transform_inputs, Bayesian7.py:178-188)."""
    b = np.asarray(bounds, dtype=np.float64)
    x_phys = np.asarray(x_unit, dtype=np.float64) * (b[1] - b[0]) + b[0]
    return (np.log(np.maximum(x_phys, 1e-6)) - np.asarray(x_log_mean)) / np.asarray(x_log_std)


def svgp_variance_score(tasks, Xs, min_variance=MIN_VARIANCE):
    """pred.variance.sum(dim=0) over the tasks (Bayesian7.py:670-671)."""
    return sum(svgp_predict(t, Xs, min_variance)[1] for t in tasks)
