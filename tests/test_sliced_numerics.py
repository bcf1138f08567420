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


def test_kernel_digit_geometry_is_int8_and_exact():
    """sweep_i8.cuh's I8Dig: eight 7-bit digits, or one 7-bit + six 8-bit ones, as bit fields of rint(x 2^F) + bias."""
    x = np.random.default_rng(1).uniform(-1, 1, 2000)
    x[:5] = [1.0, -1.0, 0.0, 2.0 ** -40, -(1.0 - 2.0 ** -50)]
    for S, W in ((8, 7), (7, 8)):
        dig = oz.field_digits(x, S, W)
        assert np.abs(dig[0]).max() <= 64                                           # the 7-bit top digit
        for dgt in dig[1:]:
            assert dgt.min() >= -(1 << (W - 1)) and dgt.max() <= (1 << (W - 1)) - 1   # signed W-bit: fits int8
        rec = dig[0] * 2.0 ** -6 + sum(dgt * 2.0 ** (-6 - W * (s + 1)) for s, dgt in enumerate(dig[1:]))
        assert np.abs(rec - x).max() <= 2.0 ** (-(6 + W * (S - 1)) - 1) * 1.0000001   # half a unit of the last place


def test_seven_wide_slices_track_eight_narrow_ones_on_reference_data():
    """The AUTO contraction: 7 slices with 8-bit lower digits (28 products) against 8 slices of 7 bits (36 products) and the
    old 7 x 7-bit form on the reference's CSV rows: a few times the 8-slice error, two orders below the 7 x 7-bit one, and
    INT32 accumulators far from overflow."""
    Li, Ks, kss, var_true = _problem()
    e64 = _rel(Li @ Ks, kss, var_true)
    U87, n87, i87 = oz.sliced_matmul_fields(Li, Ks, 8, 7)
    U78, n78, i78 = oz.sliced_matmul_fields(Li, Ks, 7, 8)
    U77, _, _ = oz.sliced_matmul_fields(Li, Ks, 7, 7)
    assert (n87, n78) == (36, 28) and max(i87, i78) < 2 ** 31
    e87, e78, e77 = (_rel(U, kss, var_true) for U in (U87, U78, U77))
    assert e87 <= 1e-9 and e78 <= 2e-9                              # both inside the bar next to clustered rows (sigma^2 ~ 3e-5)
    assert e78 <= 8 * max(e87, e64) and e77 >= 20 * e78             # ~1 bit behind 8 x 7; the 7 x 7-bit form is ~6 bits behind


import pytest  # noqa: E402


@pytest.mark.parametrize("collapsed", [False, True])
def test_svgp_stacked_factor_one_pass_variance_meets_the_bar(collapsed):
    """The sliced SVGP sweep contracts [L^-1; B], B = Ls^T L^-1 (formed in FP64 at load time), with ONE sliced panel:
    var = k** + jitter - ||L^-1 k*||^2 + ||B k*||^2.  Against the extended-precision two-solve evaluation of the same formula
    (Bayesian7.py:664-682 through gpytorch's whitened strategy) on a task with candidates on top of inducing points: the FP64
    stacked form and both sliced geometries stay inside 1e-8 -- unlike the rejected k*^T W k* form, B does not square cond(K_uu)."""
    rng = np.random.default_rng(5)
    M, d, jitter = 256, 5, 1e-6
    Z = rng.standard_normal((M, d))
    Z[40:48] = Z[7] + 1e-3 * rng.standard_normal((8, d))           # a cluster: cond(K_uu) ~ 1e7
    ls, s2 = rng.uniform(0.8, 2.0, d), 1.3
    Ls = np.tril(rng.standard_normal((M, M)) * 0.05 / np.sqrt(M / 64)) + np.diag(0.3 + 0.5 * rng.random(M))
    if collapsed:                # a trained q(u): S << I, the variance next to inducing points is 1e-3 of the prior (strong cancellation)
        Ls = np.tril(rng.standard_normal((M, M)) * 5e-4) + np.diag(0.05 * (0.5 + rng.random(M)))
    Xs = np.vstack([rng.standard_normal((192, d)), Z[:32], Z[40:48] + 1e-4, Z[100:124] + 1e-2 * rng.standard_normal((24, d))])
    K = o.kernel_matrix(Z, Z, o.KERNEL_MATERN52, ls, s2) + jitter * np.eye(M)
    L = np.linalg.cholesky(K)
    Li = np.tril(sla.solve_triangular(L, np.eye(M), lower=True, check_finite=False))
    Ks = o.kernel_matrix(Z, Xs, o.KERNEL_MATERN52, ls, s2)
    kss = o.prior_variance(Xs, o.KERNEL_MATERN52, s2) + jitter
    ld = np.longdouble
    Ul = sla.solve_triangular(L, Ks, lower=True, check_finite=False).astype(ld)
    Ul = Ul + sla.solve_triangular(L, (Ks.astype(ld) - L.astype(ld) @ Ul).astype(np.float64), lower=True).astype(ld)   # one refinement step
    Wl = Ls.T.astype(ld) @ Ul
    var_true = kss.astype(ld) - np.einsum("ij,ij->j", Ul, Ul) + np.einsum("ij,ij->j", Wl, Wl)
    assert float(var_true.min()) > 0
    B = Ls.T @ Li                                                   # what svgp_build_b forms with one FP64 GEMM
    def rel(U, W):
        var = kss - np.einsum("ij,ij->j", U, U) + np.einsum("ij,ij->j", W, W)
        return float(np.abs((var - var_true) / var_true).max())
    e64 = rel(Li @ Ks, B @ Ks)
    e78 = rel(oz.sliced_matmul_fields(Li, Ks, 7, 8)[0], oz.sliced_matmul_fields(B, Ks, 7, 8)[0])
    e87 = rel(oz.sliced_matmul_fields(Li, Ks, 8, 7)[0], oz.sliced_matmul_fields(B, Ks, 8, 7)[0])
    two_pass = rel(Li @ Ks, Ls.T @ (Li @ Ks))
    assert max(e64, two_pass) <= 1e-8, (e64, two_pass)
    assert e87 <= 1e-8 and e78 <= 1e-8, (e87, e78, e64)
