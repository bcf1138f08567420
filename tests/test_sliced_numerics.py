"""CPU tier: the numerics claim behind the INT8-sliced sweep (bayesianoptimizer_b200/csrc/sweep_i8.cuh), checked with the exact
CPU emulation of tools/ozaki_feasibility.py on the reference's own data (tests/golden/csv_n512_matern.npz: rows of
results/optimization_results.csv with duplicate and boundary points) -- no GPU involved.
Signed 7-bit slices, exact integer slice products, pairs s + t >= S dropped, FP64 recombination."""
import os
import sys

import numpy as np
import scipy.linalg as sla

from conftest import ROOT, load_golden
from oracle import gp_oracle as o

sys.path.insert(0, os.path.join(ROOT, "tools"))
import ozaki_feasibility as oz  # noqa: E402


def _problem():
    g = load_golden("csv_n512_matern")
    gp = o.fit(g["X"], g["y"], int(g["kind"]), g["lengthscale"], float(g["outputscale"]), float(g["noise"]))
    n, d = g["X"].shape
    rng = np.random.default_rng(3)
    near = np.array([np.clip(g["X"][(j * 37) % n] + eps * rng.standard_normal(d), 0, 1)
                     for j, eps in enumerate(np.logspace(-1.5, -5, 64))])
    Xs = np.vstack([g["cand"], near])
    Li = np.tril(sla.solve_triangular(gp.L, np.eye(n), lower=True, check_finite=False))
    Ks = o.kernel_matrix(gp.X, Xs, gp.kind, gp.lengthscale, gp.outputscale)
    kss = o.prior_variance(Xs, gp.kind, gp.outputscale)
    Ul = Li.astype(np.longdouble) @ Ks.astype(np.longdouble)
    var_true = kss.astype(np.longdouble) - np.einsum("ij,ij->j", Ul, Ul)
    return Li, Ks, kss, var_true


def _rel(U, kss, var_true):
    var = kss - np.einsum("ij,ij->j", U, U)
    return float(np.abs((var - var_true) / var_true).max())


def test_digits_are_int8_and_reconstruct_to_the_last_place():
    x = np.random.default_rng(0).uniform(-1, 1, 4096)
    x[:4] = [1.0, -1.0, 0.0, 2.0 ** -40]
    for S in (7, 8):
        dig, res = oz.slices(x, S)
        for dgt in dig:
            assert np.array_equal(dgt, np.rint(dgt)) and np.abs(dgt).max() <= 64
        rec = sum(dgt * 2.0 ** (-6 - 7 * s) for s, dgt in enumerate(dig))
        assert np.abs(rec - x).max() <= 2.0 ** (-7 * S)          # half a unit of the last digit
        assert np.array_equal(rec + res, x)


def test_eight_slices_meet_the_variance_bar_on_reference_data_and_seven_do_not():
    Li, Ks, kss, var_true = _problem()
    assert float(var_true.min()) < 1e-4                           # candidates next to clustered training rows: sigma^2 << noise
    e64 = _rel(Li @ Ks, kss, var_true)
    err = {}
    for S in (6, 7, 8):
        U, nprod, imax = oz.sliced_matmul(Li, Ks, S, True)
        assert nprod == S * (S + 1) // 2 and imax < 2 ** 31       # what an INT32 accumulator holds
        err[S] = _rel(U, kss, var_true)
    assert err[8] <= 1e-9 and err[8] <= 4 * e64 + 1e-12           # 8 slices (AUTO): the FP64 product's own level
    assert 1e-8 < err[7] < 1e-7                                   # 7 slices miss the 1e-8 bar next to clustered rows: opt-in only
    assert err[6] > 50 * err[7] and err[7] > 50 * err[8]          # each slice buys ~7 bits
