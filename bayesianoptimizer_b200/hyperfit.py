"""Hyper-parameter refit: MAP / maximum marginal likelihood over K7's batched LML + gradient.

Stands in for ``fit_gpytorch_mll(mll)`` (optimization/Bayesian.py:92-93; optimization/Bayesian6.py:480-488).
The reference runs SciPy L-BFGS-B on the host with one closure evaluation (K build + Cholesky + backward) per
step; here R restarts advance in LOCK STEP, so every optimiser step is ONE ``bo_lml_grad_batched`` call that
factorises all R candidate hyper-parameter vectors together on the device.  The optimiser itself is a small
box-projected L-BFGS in NumPy (plumbing: R x (d+2) numbers per step).

Parameters are theta = (log lengthscale[d], log outputscale, log noise), plus log(linear variance) for the
linear + Matern kernel (no prior on it, as in the reference's un-priored LinearKernel, Bayesian6.py:471-473).  Priors follow botorch's defaults
(SURVEY.md App. A.2): ``"lognormal"`` = botorch >= 0.12 (lengthscale ~ LogNormal(sqrt2 + log(d)/2, sqrt3),
noise ~ LogNormal(-4, 1)); ``"gamma"`` = botorch <= 0.11 (lengthscale ~ Gamma(3, 6), outputscale ~ Gamma(2, 0.15),
noise ~ Gamma(1.1, 0.05)); ``None`` = plain maximum likelihood.
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Tuple

import numpy as np


def log_prior_and_grad(theta: np.ndarray, d: int, prior: Optional[str]) -> Tuple[np.ndarray, np.ndarray]:
    """Sum of log prior densities of (lengthscale, outputscale, noise) and its gradient w.r.t. the LOG parameters.
    The densities are on the natural parameters (as gpytorch registers them), so no Jacobian term is added."""
    theta = np.atleast_2d(theta)
    lp = np.zeros(theta.shape[0])
    g = np.zeros_like(theta)
    if prior is None or prior == "none":
        return lp, g
    u_ls, u_s2, u_nz = theta[:, :d], theta[:, d], theta[:, d + 1]
    if prior == "lognormal":
        mu, sg = math.sqrt(2.0) + 0.5 * math.log(d), math.sqrt(3.0)
        # log LogNormal(x) = -log x - log(sg sqrt(2 pi)) - (log x - mu)^2 / (2 sg^2)
        lp += np.sum(-u_ls - (u_ls - mu) ** 2 / (2 * sg * sg), axis=1)
        g[:, :d] += -1.0 - (u_ls - mu) / (sg * sg)
        lp += -u_nz - (u_nz + 4.0) ** 2 / 2.0
        g[:, d + 1] += -1.0 - (u_nz + 4.0)
        return lp, g
    if prior == "gamma":
        def gam(u, a, b):      # log Gamma(x; a, rate b) up to a constant, x = exp(u); d/du = (a - 1) - b x
            x = np.exp(u)
            return (a - 1.0) * u - b * x, (a - 1.0) - b * x
        v, dv = gam(u_ls, 3.0, 6.0); lp += v.sum(axis=1); g[:, :d] += dv
        v, dv = gam(u_s2, 2.0, 0.15); lp += v; g[:, d] += dv
        v, dv = gam(u_nz, 1.1, 0.05); lp += v; g[:, d + 1] += dv
        return lp, g
    raise ValueError(f"unknown prior {prior!r}")


def lbfgs_lockstep(evaluate: Callable[[np.ndarray], Tuple[np.ndarray, np.ndarray]], theta0: np.ndarray,
                   lo: np.ndarray, hi: np.ndarray, maxiter: int = 50, history: int = 8, gtol: float = 1e-5,
                   ftol: float = 1e-9, fscale: float = 1.0, memory: Optional[dict] = None):
    """Maximise F over the box [lo, hi] from R starts at once.  ``evaluate(thetas[R,p]) -> (F[R], G[R,p])``;
    a failed evaluation returns F = -inf.  Returns (theta[R,p], F[R], n_evaluations).

    ``fscale`` is the natural scale of F (the number of observations for an un-normalised log marginal likelihood --
    the reference optimises MLL / n): the first, curvature-free step is at most 10 / fscale long per
    unit gradient.  Without it a warm start (|g| ~ 1 next to an optimum whose curvature is ~ n) overshoots by three
    orders of magnitude and spends 8 batched evaluations backtracking.

    ``memory``: a dict the caller keeps between calls on a slowly changing objective (the BO loop's warm refit: one more
    observation per iteration).  The curvature pairs (s, y) of the previous call seed this one, so the first step is already a
    quasi-Newton step instead of a short steepest-ascent one, and the pairs of this call are left in it for the next."""
    x = np.clip(np.array(theta0, dtype=np.float64, copy=True), lo, hi)
    R, p = x.shape
    f, g = evaluate(x)
    f = np.where(np.isfinite(f), f, -np.inf)
    nev = 1
    S, Y = [], []                      # each [R, p]
    if memory is not None and memory.get("shape") == (R, p):
        S, Y = [a.copy() for a in memory["S"]], [a.copy() for a in memory["Y"]]
    active = np.isfinite(f)
    for _ in range(maxiter):
        if not active.any():
            break
        # projected gradient: components pushing out of the box are dropped
        fixed = ((x <= lo) & (g < 0)) | ((x >= hi) & (g > 0))       # active bounds: work in the free subspace
        pg = g.copy()
        pg[fixed] = 0.0
        # gtol is the reference's stopping rule: fit_gpytorch_mll hands SciPy's L-BFGS-B (pgtol = 1e-5) the PER-DATUM objective
        # MLL / n, i.e. |projected gradient of F| <= gtol * fscale on the un-normalised F evaluated here
        active &= np.abs(pg).max(axis=1) > gtol * max(fscale, 1.0)
        if not active.any():
            break
        # two-loop recursion (ascent direction = H * g), per restart, vectorised over R
        q = pg.copy()
        alphas = []
        for s, y in zip(reversed(S), reversed(Y)):
            sy = np.einsum("rp,rp->r", s, y)
            rho = np.where(np.abs(sy) > 1e-300, 1.0 / np.where(sy == 0, 1.0, sy), 0.0)
            a = rho * np.einsum("rp,rp->r", s, q)
            alphas.append((a, rho, s, y))
            q -= a[:, None] * y
        if S:
            s, y = S[-1], Y[-1]
            yy = np.einsum("rp,rp->r", y, y)
            gamma = np.where(yy > 0, np.einsum("rp,rp->r", s, y) / np.where(yy == 0, 1.0, yy), 1.0)
            q *= gamma[:, None]
        for a, rho, s, y in reversed(alphas):
            b = rho * np.einsum("rp,rp->r", y, q)
            q += (a - b)[:, None] * s
        dirn = q
        dirn[fixed] = 0.0
        # fall back to the projected gradient where the quasi-Newton direction is not an ascent direction
        slope = np.einsum("rp,rp->r", dirn, pg)
        bad = ~(slope > 0)
        dirn[bad] = pg[bad]
        slope = np.einsum("rp,rp->r", dirn, pg)
        t = np.ones(R)
        if not S:
            t = np.minimum(10.0 / max(fscale, 10.0), 1.0 / np.maximum(np.abs(pg).sum(axis=1), 1e-12))
        # lock-step backtracking (Armijo) -- every trial is one batched evaluation of the still-searching restarts
        x_new, f_new, g_new = x.copy(), f.copy(), g.copy()
        searching = active.copy()
        for _ls in range(12):
            if not searching.any():
                break
            idx = np.nonzero(searching)[0]
            cand = np.clip(x[idx] + t[idx, None] * dirn[idx], lo, hi)
            fc, gc = evaluate(cand)
            nev += 1
            fc = np.where(np.isfinite(fc), fc, -np.inf)
            ok = fc >= f[idx] + 1e-4 * t[idx] * slope[idx]
            acc = idx[ok]
            x_new[acc], f_new[acc], g_new[acc] = cand[ok], fc[ok], gc[ok]
            searching[acc] = False
            t[idx[~ok]] *= 0.5
        active &= ~searching               # line search failed: that restart has converged as far as it can
        s = x_new - x
        y = -(g_new - g)                   # maximisation: curvature pair of the negated objective
        moved = np.abs(s).max(axis=1) > 0
        sy = np.einsum("rp,rp->r", s, y)
        good = moved & (sy > 1e-12)
        s[~good] = 0.0
        y[~good] = 0.0
        # the two-loop above is written for ascent on F with pairs (s, -dg); keep the sign convention consistent
        S.append(s); Y.append(y)
        if len(S) > history:
            S.pop(0); Y.pop(0)
        df = f_new - f
        x, f, g = x_new, f_new, g_new
        active &= ~(moved & (np.abs(df) <= ftol * np.maximum(1.0, np.abs(f))))
    if memory is not None:
        memory.update(shape=(R, p), S=[a.copy() for a in S], Y=[a.copy() for a in Y])
    return x, f, nev


def fit_map(engine, X, y, kernel: str, theta0: np.ndarray, lo: np.ndarray, hi: np.ndarray, prior: Optional[str] = None,
            maxiter: int = 50, mean: float = 0.0, memory: Optional[dict] = None):
    """Lock-step multi-restart MAP fit.  Returns (best theta, best objective, all thetas, all objectives, n_evals)."""
    d = int(X.shape[1])            # theta = log lengthscale[d], log outputscale, log noise [, log linear variance]

    def evaluate(th):
        lml, grad, status = engine.lml_grad_batched(X, y, th, kernel, mean)
        lml = np.asarray(lml, dtype=np.float64).copy()
        grad = np.asarray(grad, dtype=np.float64).copy()
        bad = np.asarray(status) != 0
        lp, lg = log_prior_and_grad(th, d, prior)
        F = lml + lp
        G = grad + lg
        F[bad] = -np.inf
        G[bad] = 0.0
        return F, G

    th, F, nev = lbfgs_lockstep(evaluate, theta0, lo, hi, maxiter=maxiter, fscale=float(X.shape[0]), memory=memory)
    best = int(np.argmax(F))
    return th[best], float(F[best]), th, F, nev
