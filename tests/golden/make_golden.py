"""Mint the golden fixtures from the reference's own data files with the CPU oracle.

Run in the build container (needs /root/reference, read-only):
    python tests/golden/make_golden.py
Inputs: rows [0:n] of /root/reference/results/optimization_results.csv -- the file BASELINE config 1
names -- normalised with the config/config.py:2-20 box (order n, eta, sigma_y, width, height, as
scripts/run_optimization.py:107-113 assembles it); objective = row mean of x_01..x_08
(optimization/Bayesian.py:140), standardised (botorch Standardize).  The reference holds no
known-answer vectors for this path, so these fixtures pin the ORACLE's output on
the reference's data; the GPU tests and the oracle regression tests both compare against them.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gp_oracle as o  # noqa: E402

REF = "/root/reference"
BOUNDS = np.array([(0.3, 1.0), (0.001, 300.0), (0.001, 400.0), (2.0, 7.0), (2.0, 7.0)]).T   # config/config.py:2-20
OUT = os.path.dirname(os.path.abspath(__file__))


def load_csv(n):
    raw = np.loadtxt(os.path.join(REF, "results", "optimization_results.csv"), delimiter=",", skiprows=1)
    X = o.normalize(raw[:n, :5], BOUNDS)
    y = raw[:n, 5:13].mean(axis=1)
    ys, mu, sd = o.standardize(y)
    return X, ys, mu, sd


def make(name, n, kind, ls, s2, noise, n_cand=256, seed=0):
    X, ys, ymu, ysd = load_csv(n)
    gp = o.fit(X, ys, kind, ls, s2, noise, mean=0.0)
    rng = np.random.default_rng(seed)
    cand = rng.random((n_cand, 5))
    # a few candidates sitting on / next to observed points exercise the variance cancellation
    cand[:4] = X[[0, n // 3, n // 2, n - 1]]
    cand[4:8] = np.clip(X[[1, 2, 3, 4]] + 1e-4, 0.0, 1.0)
    mu, var = o.posterior(gp, cand)
    best_f = float(ys.max())
    ei = o.acquisition(mu, var, o.ACQ_EI, best_f)
    logei = o.acquisition(mu, var, o.ACQ_LOGEI, best_f)
    ucb = o.acquisition(mu, var, o.ACQ_UCB, best_f, beta=2.0)
    tv, ti = o.topk(logei, 8)
    lml, grad = o.lml_and_grad(X, ys, kind, ls, s2, noise)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        X=X, y=ys, y_mean=ymu, y_std=ysd, kind=kind, lengthscale=np.asarray(ls, dtype=np.float64),
        outputscale=s2, noise=noise, cand=cand, alpha=gp.alpha, mu=mu, var=var, best_f=best_f, ei=ei,
        logei=logei, ucb=ucb, topk_vals=tv, topk_idx=ti, lml=lml, lml_grad=grad,
        chol_diag=np.diag(gp.L).copy())
    print(name, "n", n, "var range", var.min(), var.max(), "ei max", ei.max(), "lml", lml)


def make_cache(n=600):
    """Physical parameter rows + 8 displacements: the cached objective of BASELINE config 1 (float32 outputs)."""
    raw = np.loadtxt(os.path.join(REF, "results", "optimization_results.csv"), delimiter=",", skiprows=1)
    np.savez_compressed(os.path.join(OUT, "csv_cache_rows.npz"), params=raw[:n, :5], outputs=raw[:n, 5:13].astype(np.float32))
    print("csv_cache_rows", n)


if __name__ == "__main__":
    make_cache()
    LS = (0.5, 0.4, 0.6, 0.8, 0.7)
    make("csv_n64_matern", 64, o.KERNEL_MATERN52, LS, 1.3, 1e-3)
    make("csv_n512_matern", 512, o.KERNEL_MATERN52, LS, 1.3, 1e-3)
    make("csv_n512_rbf", 512, o.KERNEL_RBF, LS, 1.0, 1e-3)
    make("csv_n3000_matern", 3000, o.KERNEL_MATERN52, LS, 1.3, 1e-3)
