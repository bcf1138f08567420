import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["csv_n64_matern", "csv_n512_matern", "csv_n512_rbf", "csv_n3000_matern"]

# tolerances of BASELINE.json:north_star
RTOL_POST = 1e-8      # posterior mean / variance, relative
ATOL_MEAN = 1e-8      # abs floor on the mean: 1e-8 * std(y), y is standardised
RTOL_ACQ = 1e-6       # acquisition values, relative (LogEI: absolute in log space)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    g = load_golden(request.param)
    g["name"] = request.param
    return g


def synth_problem(n, d, seed_x, seed_y):
    """SURVEY.md section 8d synthetic inputs: X ~ U[0,1]^(n x d), y = sum sin(3 x_j) + 0.05 N(0,1), standardised."""
    X = np.random.default_rng(seed_x).random((n, d))
    y = np.sin(3.0 * X).sum(axis=1) + 0.05 * np.random.default_rng(seed_y).standard_normal(n)
    if n > 1:
        y = (y - y.mean()) / y.std(ddof=1)
    return X, y


def assert_posterior_close(mu, var, mu_ref, var_ref, var_abs=0.0):
    """var_abs: absolute floor on the variance tolerance, used only where the test says why (sigma^2 = s2 - ||u||^2 at
    gpytorch's 1e-6 clamp with n = 3000: the two O(s2) terms carry sqrt(n) 2^-53 s2 ~ 1e-14 of summation rounding each in
    ANY FP64 evaluation, the oracle's included, which is 1e-8 of such a variance)."""
    mu, var, mu_ref, var_ref = (np.asarray(a, dtype=np.float64) for a in (mu, var, mu_ref, var_ref))
    err_mu = np.abs(mu - mu_ref) / (RTOL_POST * np.abs(mu_ref) + ATOL_MEAN)
    assert err_mu.max() <= 1.0, f"mean off by {err_mu.max():.3g}x tolerance at {err_mu.argmax()}"
    err_var = np.abs(var - var_ref) / (RTOL_POST * np.abs(var_ref) + var_abs)
    assert err_var.max() <= 1.0, f"variance off by {err_var.max():.3g}x tolerance at {err_var.argmax()}"


def assert_acq_close(kind, val, ref):
    val, ref = np.asarray(val, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    if kind == "logei":
        err = np.abs(val - ref) / RTOL_ACQ                      # relative EI error == absolute LogEI error
    else:
        err = np.abs(val - ref) / (RTOL_ACQ * np.abs(ref) + 1e-300)
    assert err.max() <= 1.0, f"{kind} off by {err.max():.3g}x tolerance at {err.argmax()}"


def assert_ei_close_conditioned(kind, val, ref, mu_ref, var_ref, best_f):
    """EI / LogEI deep in the tail are ill-conditioned: d log EI / d u ~ -u with u = (mu - best_f) / sigma, so the relative
    error of EI (= the absolute error of LogEI) inherits the 1e-8 relative tolerance of mean and sigma amplified by
    1 + |u| + u^2 (same rule as tools/fuzz_parity.py)."""
    val, ref = np.asarray(val, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    u = (np.asarray(mu_ref) - best_f) / np.sqrt(np.asarray(var_ref))
    tol = RTOL_ACQ + 3.0 * RTOL_POST * (1.0 + np.abs(u) + u * u)
    err = np.abs(val - ref) / (tol if kind == "logei" else tol * np.abs(ref) + 1e-300)
    assert err.max() <= 1.0, f"{kind} off by {err.max():.3g}x conditioned tolerance at {err.argmax()} (u = {u[err.argmax()]:.3g})"


def refdata_pool(X, total=20_000, near=2_400, seed=7, near_min=1e-5):
    """Explicit candidate pool on the reference's CSV rows: `total - near` uniform points plus `near` points 1e-2 .. near_min
    away from training rows -- every 7th row plus the duplicated rows {12, 20}, {17, 50} of results/optimization_results.csv
    (sigma^2 down to ~1e-6 of the prior variance: where a variance contraction loses digits)."""
    n, d = X.shape
    rng = np.random.default_rng(seed)
    rows = np.concatenate([[12, 20, 17, 50], np.arange(0, n, 7)]) % n
    eps = np.logspace(-2, np.log10(near_min), near)
    pts = np.clip(X[rows[np.arange(near) % len(rows)]] + eps[:, None] * rng.standard_normal((near, d)), 0.0, 1.0)
    pool = np.vstack([rng.random((total - near, d)), pts])
    return pool[rng.permutation(total)]
