"""Slicing accuracy on the reference's own data (tests/golden fixtures: rows of results/optimization_results.csv with duplicate,
clustered and boundary points): relative error of sigma^2 for the FP64 product, 7 and 8 slices against a longdouble evaluation,
on the golden candidates plus candidates 3e-2 .. 1e-5 away from training rows.  CPU only; python tools/ozaki_golden.py"""
import os, sys, time, numpy as np, scipy.linalg as sla
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import gp_oracle as o
import ozaki_feasibility as oz
from conftest import load_golden
for name, noise_override in (("csv_n3000_matern", None), ("csv_n3000_matern", 1e-4), ("csv_n512_rbf", None)):
    g = load_golden(name)
    noise = float(g["noise"]) if noise_override is None else noise_override
    t0 = time.time()
    gp = o.fit(g["X"], g["y"], int(g["kind"]), g["lengthscale"], float(g["outputscale"]), noise)
    n, d = g["X"].shape
    rng = np.random.default_rng(3)
    near = np.array([np.clip(g["X"][(j * 37) % n] + eps * rng.standard_normal(d), 0, 1) for j, eps in enumerate(np.logspace(-1.5, -5, 64))])
    Xs = np.vstack([g["cand"][:64], near])
    Li = np.tril(sla.solve_triangular(gp.L, np.eye(n), lower=True, check_finite=False))
    Ks = o.kernel_matrix(gp.X, Xs, gp.kind, gp.lengthscale, gp.outputscale)
    kss = o.prior_variance(Xs, gp.kind, gp.outputscale)
    Ul = Li.astype(np.longdouble) @ Ks.astype(np.longdouble)
    vt = kss.astype(np.longdouble) - np.einsum("ij,ij->j", Ul, Ul)
    ok = np.asarray(vt > 1e-6)
    def rel(U):
        v = kss - np.einsum("ij,ij->j", U, U)
        return float(np.abs((v - vt) / vt)[ok].max())
    out = [f"{name} noise={noise:g} ratio={noise/float(g['outputscale']):.1e} n={n} max|Li|={np.abs(Li).max():.3g} min sigma^2={float(vt[ok].min()):.2e}: fp64 {rel(Li @ Ks):.1e}"]
    for S in (7, 8):
        U, _, imax = oz.sliced_matmul(Li, Ks, S, True)
        out.append(f"S={S} {rel(U):.1e} (acc 2^{np.log2(imax):.1f})")
    print("  ".join(out), f"[{time.time()-t0:.0f} s]", flush=True)
