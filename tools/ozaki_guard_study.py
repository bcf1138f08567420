"""CPU study for DESIGN section 7 (2b): a per-candidate estimate of the 7-slice error of ||u||^2 that the drain could compute
for free, so that only candidates whose sigma^2 is too small for it need the 8-slice / FP64 value.

    delta(||u||^2) = 2 sum_i u_i delta_i,   delta_i ~ 2^-49 rowscale_i eb sqrt(i + 1) c      (independent slicing errors)
    est_c = kappa 2^-49 eb sqrt( sum_i (rowscale_i u_i)^2 (i + 1) )

The drain already holds u_i and rowscale_i; one extra FMA per element accumulates the sum.  Reported: the kappa that
covers every candidate of the reference-data fixtures and the C3-shaped synthetic problem, and how many candidates a rule
"flag if sigma^2 < 1e9 est" (7-slice error above 1e-9 relative) would send to the exact path: random pool points vs points
next to training rows.   python tools/ozaki_guard_study.py"""
import os, sys, numpy as np, scipy.linalg as sla
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import gp_oracle as o
import ozaki_feasibility as oz
from conftest import load_golden, synth_problem


def problem(name):
    if name == "c3_2048":
        X, y = synth_problem(2048, 8, 4, 5)
        return X, y, o.KERNEL_MATERN52, np.full(8, 0.7), 1.0, 1e-3
    g = load_golden(name)
    return g["X"], g["y"], int(g["kind"]), g["lengthscale"], float(g["outputscale"]), float(g["noise"])


for name in ("csv_n512_matern", "csv_n512_rbf", "csv_n3000_matern", "c3_2048"):
    X, y, kind, ls, s2, noise = problem(name)
    n, d = X.shape
    gp = o.fit(X, y, kind, ls, s2, noise)
    rng = np.random.default_rng(3)
    far = rng.random((192, d))
    near = np.array([np.clip(X[(j * 37) % n] + eps * rng.standard_normal(d), 0, 1) for j, eps in enumerate(np.logspace(-1.5, -5, 64))])
    Xs = np.vstack([far, near])
    Li = np.tril(sla.solve_triangular(gp.L, np.eye(n), lower=True, check_finite=False))
    Ks = o.kernel_matrix(gp.X, Xs, gp.kind, gp.lengthscale, gp.outputscale)
    kss = o.prior_variance(Xs, gp.kind, gp.outputscale)
    Ul = Li.astype(np.longdouble) @ Ks.astype(np.longdouble)
    ss_true = np.einsum("ij,ij->j", Ul, Ul)
    var_true = (kss.astype(np.longdouble) - ss_true).astype(np.float64)
    U7, _, _ = oz.sliced_matmul(Li, Ks, 7, True)
    err = np.abs((np.einsum("ij,ij->j", U7, U7) - ss_true).astype(np.float64))
    ra = oz.pow2_scale(Li, 1).reshape(-1)
    eb = float(oz.pow2_scale(np.array([[s2]]), 0).reshape(-1)[0])
    w = (np.arange(n) + 1.0)
    est1 = 2.0 ** -49 * eb * np.sqrt(np.einsum("i,ij->j", (ra ** 2) * w, U7 * U7))
    kappa = float((err / est1).max())
    k_use = 4.0
    flag = var_true < 1e9 * k_use * est1
    rel = err / var_true
    print(f"{name:<18s} n={n}: max err/est = {kappa:.2f} (median {np.median(err / est1):.2f}); with kappa = {k_use:g}: flagged {int(flag[:192].sum())}/192 random, "
          f"{int(flag[192:].sum())}/64 near-training; worst 7-slice rel error among UNflagged = {rel[~flag].max() if (~flag).any() else 0:.1e}, among flagged = {rel[flag].max() if flag.any() else 0:.1e}",
          flush=True)
