"""CPU tier: the oracle against the golden fixtures and against independently written implementations
(scikit-learn GPR, scipy.stats.norm, finite differences, torch's SobolEngine)."""
import numpy as np
import pytest
import scipy.stats as sst

from conftest import assert_acq_close, assert_posterior_close, load_golden, synth_problem
from oracle import gp_oracle as o


def test_oracle_reproduces_golden(golden):
    g = golden
    gp = o.fit(g["X"], g["y"], int(g["kind"]), g["lengthscale"], float(g["outputscale"]), float(g["noise"]))
    np.testing.assert_allclose(gp.alpha, g["alpha"], rtol=1e-9, atol=1e-9 * np.abs(g["alpha"]).max())
    np.testing.assert_allclose(np.diag(gp.L), g["chol_diag"], rtol=1e-11)
    mu, var = o.posterior(gp, g["cand"])
    assert_posterior_close(mu, var, g["mu"], g["var"])
    bf = float(g["best_f"])
    assert_acq_close("ei", o.acquisition(mu, var, o.ACQ_EI, bf), g["ei"])
    assert_acq_close("logei", o.acquisition(mu, var, o.ACQ_LOGEI, bf), g["logei"])
    assert_acq_close("ucb", o.acquisition(mu, var, o.ACQ_UCB, bf, 2.0), g["ucb"])
    tv, ti = o.topk(o.acquisition(mu, var, o.ACQ_LOGEI, bf), 8)
    assert ti.tolist() == g["topk_idx"].tolist()
    lml, grad = o.lml_and_grad(g["X"], g["y"], int(g["kind"]), g["lengthscale"], float(g["outputscale"]), float(g["noise"]))
    assert abs(lml - float(g["lml"])) <= 1e-9 * abs(float(g["lml"]))
    np.testing.assert_allclose(grad, g["lml_grad"], rtol=1e-7, atol=1e-7)


@pytest.mark.parametrize("name", ["csv_n64_matern", "csv_n512_matern"])
def test_oracle_matches_sklearn_gpr(name):
    """Independent exact-GP implementation (direct-difference distances, cho_solve, solve_triangular)."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern
    g = load_golden(name)
    kern = ConstantKernel(float(g["outputscale"])) * Matern(g["lengthscale"], nu=2.5)
    gpr = GaussianProcessRegressor(kernel=kern, alpha=float(g["noise"]), optimizer=None).fit(g["X"], g["y"])
    m, s = gpr.predict(g["cand"], return_std=True)
    far = g["var"] > 1e-2          # sklearn loses digits in std**2 near the data; compare where it is well conditioned
    np.testing.assert_allclose(g["mu"], m, rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(g["var"][far], (s ** 2)[far], rtol=1e-7)
    assert abs(gpr.log_marginal_likelihood_value_ - float(g["lml"])) <= 1e-8 * abs(float(g["lml"]))


def test_rbf_matches_sklearn_gpr():
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF
    g = load_golden("csv_n512_rbf")
    gpr = GaussianProcessRegressor(kernel=RBF(g["lengthscale"]), alpha=float(g["noise"]), optimizer=None).fit(g["X"], g["y"])
    m = gpr.predict(g["cand"])
    np.testing.assert_allclose(g["mu"], m, rtol=1e-6, atol=1e-7)


def test_ei_ucb_closed_forms_against_scipy():
    rng = np.random.default_rng(0)
    mu = rng.standard_normal(2000)
    var = rng.random(2000) * 2 + 1e-6
    sig = np.sqrt(var)
    bf = 0.3
    u = (mu - bf) / sig
    ei_ref = sig * (sst.norm.pdf(u) + u * sst.norm.cdf(u))
    ei = o.acquisition(mu, var, o.ACQ_EI, bf)
    ok = ei_ref > 1e-12
    np.testing.assert_allclose(ei[ok], ei_ref[ok], rtol=1e-9)
    np.testing.assert_allclose(o.acquisition(mu, var, o.ACQ_UCB, bf, 4.0), mu + 2.0 * sig, rtol=1e-14)
    # LogEI agrees with log(EI) where EI is representable, and stays finite far in the tail
    le = o.acquisition(mu, var, o.ACQ_LOGEI, bf)
    np.testing.assert_allclose(le[ok], np.log(ei_ref[ok]), rtol=1e-8, atol=1e-8)
    tail = o.acquisition(np.array([-40.0, -300.0]), np.array([1.0, 1.0]), o.ACQ_LOGEI, 0.0)
    assert np.all(np.isfinite(tail)) and tail[1] < tail[0] < -700
    # asymptote: log h(u) ~ -u^2/2 - log(sqrt(2 pi)) - 2 log|u|
    assert abs(tail[1] - (-0.5 * 300.0 ** 2 - 0.5 * np.log(2 * np.pi) - 2 * np.log(300.0))) < 1e-3


def test_lml_gradient_against_finite_differences():
    X, y = synth_problem(120, 4, 11, 12)
    for kind in (o.KERNEL_MATERN52, o.KERNEL_RBF):
        theta = np.log(np.array([0.5, 0.4, 0.6, 0.8, 1.3, 1e-2]))

        def f(t):
            return o.lml_and_grad(X, y, kind, np.exp(t[:4]), np.exp(t[4]), np.exp(t[5]))[0]

        _, grad = o.lml_and_grad(X, y, kind, np.exp(theta[:4]), np.exp(theta[4]), np.exp(theta[5]))
        fd = np.array([(f(theta + 1e-6 * e) - f(theta - 1e-6 * e)) / 2e-6 for e in np.eye(6)])
        np.testing.assert_allclose(grad, fd, rtol=2e-6, atol=1e-6)


def test_acquisition_gradient_against_finite_differences():
    X, y = synth_problem(150, 5, 1, 2)
    rng = np.random.default_rng(5)
    for kind in (o.KERNEL_MATERN52, o.KERNEL_RBF, o.KERNEL_LINEAR_MATERN52):      # kind 2: Bayesian6.py:470-478
        gp = o.fit(X, y, kind, [0.5, 0.4, 0.6, 0.8, 0.7], 1.3, 1e-3, linear_variance=0.37 if kind == o.KERNEL_LINEAR_MATERN52 else 0.0)
        x0 = rng.random(5)
        for ak in (o.ACQ_EI, o.ACQ_LOGEI, o.ACQ_UCB, o.ACQ_VAR, o.ACQ_MEAN):
            _, gd = o.acquisition_with_grad(gp, x0, ak, 1.0, 2.0)
            fd = np.array([(o.acquisition_with_grad(gp, x0 + 1e-6 * e, ak, 1.0, 2.0)[0]
                            - o.acquisition_with_grad(gp, x0 - 1e-6 * e, ak, 1.0, 2.0)[0]) / 2e-6 for e in np.eye(5)])
            np.testing.assert_allclose(gd, fd, rtol=2e-5, atol=1e-9)


def test_row_append_identities():
    """believer append keeps alpha' = [alpha; 0] and the mean, and shrinks the variance."""
    X, y = synth_problem(100, 3, 3, 4)
    gp = o.fit(X, y, o.KERNEL_MATERN52, 0.6, 1.0, 1e-3)
    xs = np.random.default_rng(9).random((50, 3))
    mu0, var0 = o.posterior(gp, xs, min_variance=0.0)
    xnew = np.array([0.31, 0.62, 0.17])
    gp2 = o.append_point(gp, xnew)
    np.testing.assert_allclose(gp2.alpha[:-1], gp.alpha, rtol=1e-6, atol=1e-8)
    assert abs(gp2.alpha[-1]) < 1e-8
    mu1, var1 = o.posterior(gp2, xs, min_variance=0.0)
    np.testing.assert_allclose(mu1, mu0, rtol=1e-8, atol=1e-9)
    assert np.all(var1 <= var0 + 1e-12)
    # an observed append equals a refit on the extended data
    gp3 = o.append_point(gp, xnew, 0.5)
    ref = o.fit(np.vstack([X, xnew]), np.concatenate([y, [0.5]]), o.KERNEL_MATERN52, 0.6, 1.0, 1e-3)
    np.testing.assert_allclose(gp3.alpha, ref.alpha, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(gp3.L, ref.L, rtol=1e-9, atol=1e-12)


def test_not_positive_definite_and_jitter_retry():
    """Exact duplicate rows + zero noise -> Cholesky failure with a pivot; jitter 1e-2 rescues it
    (the convention of optimization/Bayesian6.py:482-488)."""
    g = load_golden("csv_n3000_matern")
    X, y = g["X"][:400], g["y"][:400]
    with pytest.raises(o.NotPositiveDefinite) as ei:
        o.fit(X, y, o.KERNEL_RBF, 2.0, 1.0, 0.0)
    assert 1 <= ei.value.pivot <= 400
    o.fit(X, y, o.KERNEL_RBF, 2.0, 1.0, 0.0, jitter=1e-2)


def test_topk_order_ties_and_nan():
    v = np.array([1.0, 3.0, np.nan, 3.0, 2.0, 3.0])
    tv, ti = o.topk(v, 4, first_index=100)
    assert ti.tolist() == [101, 103, 105, 104]
    assert tv.tolist() == [3.0, 3.0, 3.0, 2.0]
    mv, mi = o.merge_topk([tv[:2], np.array([3.0, 0.5])], [ti[:2], np.array([7, 9])], 3)
    assert mi.tolist() == [7, 101, 103]


def test_sobol_restatement_matches_torch():
    import torch
    for d, seed in ((8, 6), (5, 3), (16, 1)):
        eng = torch.quasirandom.SobolEngine(d, scramble=True, seed=seed)
        st, sh = eng.sobolstate.numpy().copy(), eng.shift.numpy().copy()
        ref = eng.draw(3000, dtype=torch.float64).numpy()
        assert np.array_equal(o.sobol_points(st, sh, 0, 3000), ref)
        assert np.array_equal(o.sobol_points(st, sh, 1234, 100), ref[1234:1334])


def test_transforms():
    b = np.array([(0.3, 1.0), (0.001, 300.0)]).T
    X = np.array([[0.3, 300.0], [0.65, 150.0005]])
    U = o.normalize(X, b)
    np.testing.assert_allclose(U, [[0.0, 1.0], [0.5, 0.5]], atol=1e-12)
    np.testing.assert_allclose(o.unnormalize(U, b), X, rtol=1e-14)
    ys, m, s = o.standardize(np.array([1.0, 2.0, 3.0, 4.0]))
    assert abs(m - 2.5) < 1e-15 and abs(s - np.std([1, 2, 3, 4], ddof=1)) < 1e-15
    assert abs(ys.mean()) < 1e-15


# ---- N4: ScaleKernel(Linear + Matern-5/2), outputs sharing one kernel matrix -----------------------
def _lin_problem(n=140, d=4):
    X, y = synth_problem(n, d, 21, 22)
    y = y + 0.8 * (X @ np.array([1.0, -0.5, 0.25, 0.7])[:d])          # a linear trend for the linear part to pick up
    return X, (y - y.mean()) / y.std(ddof=1)


def test_linear_matern_matches_sklearn_gpr():
    """s2 (v <x,x'> + matern52) against sklearn's ConstantKernel * (ConstantKernel * DotProduct + Matern)."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import ConstantKernel, DotProduct, Matern
    X, y = _lin_problem()
    ls, s2, v, noise = [0.5, 0.4, 0.6, 0.8], 1.3, 0.37, 1e-3
    gp = o.fit(X, y, o.KERNEL_LINEAR_MATERN52, ls, s2, noise, linear_variance=v)
    kern = ConstantKernel(s2) * (ConstantKernel(v) * DotProduct(sigma_0=1e-150) + Matern(ls, nu=2.5))
    gpr = GaussianProcessRegressor(kernel=kern, alpha=noise, optimizer=None).fit(X, y)
    xs = np.random.default_rng(3).random((300, 4)) * 1.2 - 0.1
    m, s = gpr.predict(xs, return_std=True)
    mu, var = o.posterior(gp, xs)
    np.testing.assert_allclose(mu, m, rtol=1e-7, atol=1e-8)
    far = var > 1e-2
    np.testing.assert_allclose(var[far], (s ** 2)[far], rtol=1e-7)
    lml, _ = o.lml_and_grad(X, y, o.KERNEL_LINEAR_MATERN52, ls, s2, noise, linear_variance=v)
    assert abs(gpr.log_marginal_likelihood_value_ - lml) <= 1e-8 * abs(lml)


def test_linear_matern_lml_gradient_against_finite_differences():
    X, y = _lin_problem(110, 3)
    theta = np.log(np.array([0.5, 0.4, 0.6, 1.3, 1e-2, 0.37]))

    def f(t):
        return o.lml_and_grad(X, y, o.KERNEL_LINEAR_MATERN52, np.exp(t[:3]), np.exp(t[3]), np.exp(t[4]), linear_variance=np.exp(t[5]))[0]

    _, grad = o.lml_and_grad(X, y, o.KERNEL_LINEAR_MATERN52, np.exp(theta[:3]), np.exp(theta[3]), np.exp(theta[4]),
                             linear_variance=np.exp(theta[5]))
    assert grad.shape == (6,)
    fd = np.array([(f(theta + 1e-6 * e) - f(theta - 1e-6 * e)) / 2e-6 for e in np.eye(6)])
    np.testing.assert_allclose(grad, fd, rtol=2e-6, atol=1e-6)


def test_linear_matern_append_equals_refit():
    X, y = _lin_problem(90, 4)
    kw = dict(kind=o.KERNEL_LINEAR_MATERN52, lengthscale=0.6, outputscale=1.1, noise=1e-3, linear_variance=0.5)
    gp = o.fit(X, y, **kw)
    xnew = np.array([0.9, 0.1, 0.5, 0.3])
    gp2 = o.append_point(gp, xnew, 0.25)
    ref = o.fit(np.vstack([X, xnew]), np.concatenate([y, [0.25]]), **kw)
    np.testing.assert_allclose(gp2.L, ref.L, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(gp2.alpha, ref.alpha, rtol=1e-8, atol=1e-10)
    xs = np.random.default_rng(1).random((40, 4))
    mu0, _ = o.posterior(gp, xs)
    mu1, _ = o.posterior(o.append_point(gp, xnew), xs)                 # believer: mean unchanged
    np.testing.assert_allclose(mu1, mu0, rtol=1e-8, atol=1e-9)


def test_posterior_multi_equals_separate_fits():
    X, _ = _lin_problem(100, 4)
    rng = np.random.default_rng(5)
    Y = np.stack([np.sin(3 * X).sum(1), np.cos(2 * X).sum(1), X @ rng.random(4)], axis=1) + 0.01 * rng.standard_normal((100, 3))
    means = np.array([0.1, -0.2, 0.0])
    xs = rng.random((60, 4))
    for kind, v in ((o.KERNEL_MATERN52, 0.0), (o.KERNEL_LINEAR_MATERN52, 0.4)):
        gp = o.fit(X, Y[:, 0], kind, 0.5, 1.0, 1e-3, mean=means[0], linear_variance=v)
        mu, var = o.posterior_multi(gp, Y, xs, means)
        for t in range(3):
            gpt = o.fit(X, Y[:, t], kind, 0.5, 1.0, 1e-3, mean=means[t], linear_variance=v)
            mt, vt = o.posterior(gpt, xs)
            np.testing.assert_allclose(mu[:, t], mt, rtol=1e-10, atol=1e-12)
            np.testing.assert_allclose(var, vt, rtol=1e-12)


def test_log_standardize_round_trip_and_lognormal_mean():
    """Output transform of optimization/Bayesian6.py:427-443 and its back-transform (:631-633, :703-707)."""
    rng = np.random.default_rng(0)
    Y = rng.random((50, 8)) * 10 + 0.05
    tr = o.LogStandardize.fit(Y)
    assert tr.shift == pytest.approx(max(1e-12, np.abs(Y).max() * 1e-6))
    Z = tr.forward(Y)
    np.testing.assert_allclose(Z.mean(axis=0), 0.0, atol=1e-12)
    np.testing.assert_allclose(Z.std(axis=0, ddof=1), 1.0, rtol=1e-12)
    np.testing.assert_allclose(tr.inverse_mean(Z, np.zeros_like(Z)), Y, rtol=1e-12)       # zero variance: exact inverse
    lm = tr.inverse_mean(Z, np.full_like(Z, 0.3))
    np.testing.assert_allclose(lm, np.exp(np.log(Y + tr.shift) + 0.5 * 0.3 * tr.std ** 2) - tr.shift, rtol=1e-12)
    # non-positive data: the shift lifts the minimum above zero
    Yn = Y - 3.0
    tn = o.LogStandardize.fit(Yn)
    assert tn.shift == pytest.approx(-Yn.min() + max(1e-12, np.abs(Yn).max() * 1e-6))
    assert np.all(np.isfinite(tn.forward(Yn)))


# ---- N2: SVGP predictive (whitened variational strategy) ---------------------------------------------------------
def _svgp_task(M=60, d=4, seed=0, kind=o.KERNEL_LINEAR_MATERN52):
    rng = np.random.default_rng(seed)
    Z = rng.standard_normal((M, d))
    A = rng.standard_normal((M, M)) * 0.05
    Ls = np.tril(A) + np.diag(0.3 + 0.5 * rng.random(M))
    return o.SVGPTask(Z, kind, rng.uniform(0.5, 1.5, d), 1.3, 0.2 if kind == o.KERNEL_LINEAR_MATERN52 else 0.0,
                      0.1, 2e-3, 1e-6, rng.standard_normal(M), Ls)


def test_svgp_predictive_matches_unwhitened_formula():
    """Independent route: q(u) = N(L m, L S L^T) -> mean = c + k^T K^-1 mu_u, var = k** - k^T K^-1 (K - Sigma_u) K^-1 k."""
    for kind in (o.KERNEL_LINEAR_MATERN52, o.KERNEL_MATERN52, o.KERNEL_RBF):
        t = _svgp_task(kind=kind)
        xs = np.random.default_rng(1).standard_normal((200, 4))
        mean, var = o.svgp_predict(t, xs, min_variance=0.0)
        K = o.kernel_matrix(t.Z, t.Z, t.kind, t.lengthscale, t.outputscale, t.linear_variance) + t.jitter * np.eye(len(t.Z))
        L = np.linalg.cholesky(K)
        mu_u, Sig_u = L @ t.m, L @ np.tril(t.Ls) @ np.tril(t.Ls).T @ L.T
        ks = o.kernel_matrix(t.Z, xs, t.kind, t.lengthscale, t.outputscale, t.linear_variance)
        A = np.linalg.solve(K, ks)
        np.testing.assert_allclose(mean, t.mean + A.T @ mu_u, rtol=1e-8, atol=1e-9)
        ref = (o.prior_variance(xs, t.kind, t.outputscale, t.linear_variance) + t.jitter
               - np.einsum("ij,ij->j", ks, A) + np.einsum("ij,ij->j", A, Sig_u @ A) + t.noise)
        np.testing.assert_allclose(var, ref, rtol=1e-7, atol=1e-9)


def test_svgp_with_optimal_variational_posterior_equals_exact_gp():
    """Z = X and the optimal Gaussian q(u) (Titsias): the SVGP predictive is the exact GP posterior."""
    X, y = synth_problem(80, 3, 5, 6)
    ls, s2, noise = np.array([0.6, 0.5, 0.7]), 1.2, 5e-2
    gp = o.fit(X, y, o.KERNEL_MATERN52, ls, s2, noise)
    K = o.kernel_matrix(X, X, o.KERNEL_MATERN52, ls, s2)
    L = np.linalg.cholesky(K)
    Sig_u = np.linalg.inv(np.linalg.inv(K) + np.eye(80) / noise)
    mu_u = Sig_u @ y / noise
    S = np.linalg.solve(L, np.linalg.solve(L, Sig_u).T)
    t = o.SVGPTask(X, o.KERNEL_MATERN52, ls, s2, 0.0, 0.0, 0.0, 0.0, np.linalg.solve(L, mu_u), np.linalg.cholesky((S + S.T) / 2))
    xs = np.random.default_rng(2).random((100, 3))
    mean, var = o.svgp_predict(t, xs)
    omu, ovar = o.posterior(gp, xs)
    np.testing.assert_allclose(mean, omu, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(var, ovar, rtol=1e-5, atol=1e-9)


def test_svgp_prior_and_transform():
    t = _svgp_task()
    t.m[:] = 0.0
    t.Ls = np.eye(len(t.Z))                                       # q(u) = prior -> predictive = prior
    xs = np.random.default_rng(3).standard_normal((50, 4))
    mean, var = o.svgp_predict(t, xs)
    np.testing.assert_allclose(mean, t.mean, atol=1e-12)
    np.testing.assert_allclose(var, o.prior_variance(xs, t.kind, t.outputscale, t.linear_variance) + t.jitter + t.noise, rtol=1e-9)
    b = np.array([[0.3, 0.001], [1.0, 300.0]])
    z = o.svgp_transform_inputs(np.array([[0.0, 0.0], [1.0, 1.0]]), b, np.array([[0.0, 1.0]]), np.array([[2.0, 0.5]]))
    np.testing.assert_allclose(z, [[np.log(0.3) / 2, (np.log(0.001) - 1) / 0.5], [0.0, (np.log(300.0) - 1) / 0.5]], rtol=1e-12)
