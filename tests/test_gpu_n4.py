"""GPU tier, SURVEY 8f N4: ScaleKernel(Linear + Matern-5/2) (optimization/Bayesian6.py:471-473, Bayesian7.py:162-166)
through fit / posterior / sweep / append / batched LML, and m outputs sharing one factorisation (bo_posterior_multi),
each against the CPU oracle through the C ABI."""
import numpy as np
import pytest

from conftest import assert_acq_close, assert_posterior_close, synth_problem
from oracle import gp_oracle as o

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

LIN = o.KERNEL_LINEAR_MATERN52


@pytest.fixture(scope="module")
def engine():
    from bayesianoptimizer_b200 import GPEngine
    eng = GPEngine(torch.device("cuda", 0))
    yield eng
    eng.close()


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _problem(n, d, seed):
    X, y = synth_problem(n, d, seed, seed + 1)
    y = y + 0.8 * (X @ np.linspace(1.0, -0.5, d))
    return X, (y - y.mean()) / y.std(ddof=1)


@pytest.mark.parametrize("n,d,N", [(300, 5, 2000), (129, 3, 100), (1000, 8, 4000), (640, 16, 300), (2500, 5, 10000)])
def test_linear_matern_posterior_and_sweep(engine, n, d, N):
    """Fused sweep (incl. the row-split path of small pools) with the per-candidate prior variance s2 (v |x|^2 + 1)."""
    X, y = _problem(n, d, 31)
    ls = np.linspace(0.4, 0.9, d)
    kw = dict(lengthscale=ls, outputscale=1.3, noise=1e-3, mean=0.05, linear_variance=0.37)
    gp = o.fit(X, y, LIN, **kw)
    engine.fit(_cuda(X), _cuda(y), "linear_matern52", ls, 1.3, 1e-3, mean=0.05, linear_variance=0.37)
    alpha, L, _ = (t.cpu().numpy() for t in engine.state())
    np.testing.assert_allclose(np.diag(L), np.diag(gp.L), rtol=1e-9)
    assert np.abs(alpha - gp.alpha).max() <= 1e-7 * np.abs(gp.alpha).max()
    xs = np.random.default_rng(7).random((N, d)) * 1.1 - 0.05
    xs[0] = X[3]
    bf = float(y.max())
    for acq, ak in (("ei", o.ACQ_EI), ("logei", o.ACQ_LOGEI), ("ucb", o.ACQ_UCB)):
        vals, idx, mu, var, av = engine.sweep(acq, bf, 2.0, candidates=_cuda(xs), topk=5, return_all=True)
        tv, ti, omu, ovar, oav = o.sweep(gp, xs, ak, bf, 2.0, k=5)
        assert_posterior_close(mu.cpu().numpy(), var.cpu().numpy(), omu, ovar)
        assert_acq_close(acq, av.cpu().numpy(), oav)
        assert idx.cpu().tolist() == ti.tolist()
    mu, var = engine.posterior(_cuda(xs[:50]))
    omu, ovar = o.posterior(gp, xs[:50])
    assert_posterior_close(mu.cpu().numpy(), var.cpu().numpy(), omu, ovar)


def test_linear_matern_reference_kernel_agrees(engine, monkeypatch):
    """Two-implementation cross-check on the device (plain-load kernel vs the TMA/DMMA kernel)."""
    X, y = _problem(400, 6, 5)
    engine.fit(_cuda(X), _cuda(y), "linear_matern52", 0.6, 1.1, 1e-3, linear_variance=0.5)
    xs = _cuda(np.random.default_rng(1).random((500, 6)))
    fast = engine.sweep("ei", 0.5, candidates=xs, topk=3, return_all=True)
    monkeypatch.setenv("BO_B200_SWEEP_IMPL", "reference")
    slow = engine.sweep("ei", 0.5, candidates=xs, topk=3, return_all=True)
    assert fast[1].tolist() == slow[1].tolist()
    assert_posterior_close(fast[2].cpu().numpy(), fast[3].cpu().numpy(), slow[2].cpu().numpy(), slow[3].cpu().numpy())


def test_linear_matern_sobol_pool_and_append(engine):
    from bayesianoptimizer_b200 import sobol_state
    d = 4
    X, y = _problem(250, d, 41)
    kw = dict(lengthscale=0.5, outputscale=1.0, noise=1e-3, linear_variance=0.8)
    gp = o.fit(X, y, LIN, **kw)
    engine.fit(_cuda(X), _cuda(y), "linear_matern52", 0.5, 1.0, 1e-3, linear_variance=0.8)
    st = sobol_state(d, 9)
    se = torch.quasirandom.SobolEngine(d, scramble=True, seed=9)
    pts = o.sobol_points(se.sobolstate.numpy(), se.shift.numpy(), 0, 3000)
    bf = float(y.max())
    for step in range(3):                                     # believer, believer, observed
        v, i = engine.sweep("logei", bf, sobol=st, count=3000, topk=1)
        tv, ti, _, _, _ = o.sweep(gp, pts, o.ACQ_LOGEI, bf, k=1)
        assert i.item() == ti[0]
        x = engine.sobol_points(st, i)[0]
        yy = None if step < 2 else 0.3
        engine.append(x, yy)
        gp = o.append_point(gp, pts[ti[0]], yy)
    xs = np.random.default_rng(2).random((100, d))
    mu, var = engine.posterior(_cuda(xs))
    omu, ovar = o.posterior(gp, xs)
    assert_posterior_close(mu.cpu().numpy(), var.cpu().numpy(), omu, ovar)
    ref = o.fit(gp.X, gp.y, LIN, **kw)
    np.testing.assert_allclose(np.diag(engine.state()[1].cpu().numpy()), np.diag(ref.L), rtol=1e-8)


@pytest.mark.parametrize("acq,ak", [("ei", o.ACQ_EI), ("logei", o.ACQ_LOGEI), ("ucb", o.ACQ_UCB), ("var", o.ACQ_VAR), ("mean", o.ACQ_MEAN)])
def test_linear_matern_acquisition_gradient_against_oracle(engine, acq, ak):
    """bo_acq_grad on the reference's explicit kernel ScaleKernel(Linear + Matern) (optimization/Bayesian6.py:470-478 -- the
    kernel its optimize_acqf call at :905-916 differentiates through autograd): values and gradients against the oracle's
    analytic formulas, which tests/test_oracle.py pins to finite differences."""
    X, y = _problem(300, 5, 3)
    ls = np.array([0.5, 0.4, 0.6, 0.8, 0.7])
    gp = o.fit(X, y, LIN, ls, 1.3, 1e-3, mean=0.05, linear_variance=0.37)
    engine.fit(_cuda(X), _cuda(y), "linear_matern52", ls, 1.3, 1e-3, mean=0.05, linear_variance=0.37)
    rng = np.random.default_rng(9)
    Q = np.vstack([rng.random((24, 5)), np.clip(X[:8] + 1e-3 * rng.standard_normal((8, 5)), 0, 1)])
    bf = float(y.max())
    val, grad = engine.acq_grad(_cuda(Q), acq, bf, 2.0)
    val, grad = val.cpu().numpy(), grad.cpu().numpy()
    for i, x in enumerate(Q):
        v, g = o.acquisition_with_grad(gp, x, ak, bf, 2.0)
        assert abs(val[i] - v) <= 1e-6 * max(abs(v), 1e-300) + (1e-6 if acq == "logei" else 0.0), (i, val[i], v)
        np.testing.assert_allclose(grad[i], g, rtol=2e-6, atol=1e-8 * (1.0 + np.abs(g).max()))


def test_linear_matern_refine_matches_scipy_lbfgsb(engine):
    """bo_refine on the linear + Matern kind vs SciPy L-BFGS-B on the oracle's acquisition from the same starts
    (the optimize_acqf call of Bayesian6.py:905-916); SURVEY App. A.7 criterion as for the stationary kinds."""
    import scipy.optimize as so
    from bayesianoptimizer_b200 import BoError
    X, y = _problem(256, 4, 5)
    gp = o.fit(X, y, LIN, 0.6, 1.0, 1e-3, linear_variance=0.5)
    engine.fit(_cuda(X), _cuda(y), "linear_matern52", 0.6, 1.0, 1e-3, linear_variance=0.5)
    bf = float(y.max())
    starts = np.random.default_rng(5).random((12, 4))
    v0, _ = engine.acq_grad(_cuda(starts), "logei", bf)
    xr, vr = engine.refine(_cuda(starts), "logei", bf, iters=150)
    xr, vr, v0 = xr.cpu().numpy(), vr.cpu().numpy(), v0.cpu().numpy()
    assert np.all(vr >= v0 - 1e-12) and np.all((xr >= 0.0) & (xr <= 1.0))
    for i in range(12):
        v, _ = o.acquisition_with_grad(gp, xr[i], o.ACQ_LOGEI, bf)
        assert abs(v - vr[i]) <= 1e-6
    ref = np.array([-so.minimize(lambda z: tuple(-np.asarray(t) for t in o.acquisition_with_grad(gp, z, o.ACQ_LOGEI, bf)), s0, jac=True,
                                 method="L-BFGS-B", bounds=[(0.0, 1.0)] * 4, options={"maxiter": 200}).fun for s0 in starts])
    assert vr.max() >= ref.max() - 1e-3 * max(1.0, abs(ref.max()))
    assert np.mean(vr >= ref - 1e-2 * np.maximum(1.0, np.abs(ref))) >= 0.75
    with pytest.raises(BoError):
        engine.fit(_cuda(X), _cuda(y), "linear_matern52", 0.5, 1.0, 1e-3, linear_variance=-1.0)


@pytest.mark.parametrize("n,d", [(300, 5), (700, 10), (130, 2)])
def test_linear_matern_batched_lml_and_gradient(engine, n, d):
    X, y = _problem(n, d, 8)
    rng = np.random.default_rng(9)
    R = 5
    th = np.concatenate([rng.uniform(np.log(0.2), np.log(3.0), (R, d)), rng.uniform(np.log(0.5), np.log(2.0), (R, 1)),
                         rng.uniform(np.log(1e-3), np.log(1e-1), (R, 1)), rng.uniform(np.log(0.05), np.log(2.0), (R, 1))], axis=1)
    lml, grad, status = engine.lml_grad_batched(_cuda(X), _cuda(y), th, "linear_matern52")
    assert status.tolist() == [0] * R and grad.shape == (R, d + 3)
    for r in range(R):
        l, g = o.lml_and_grad(X, y, LIN, np.exp(th[r, :d]), np.exp(th[r, d]), np.exp(th[r, d + 1]), linear_variance=np.exp(th[r, d + 2]))
        assert abs(lml[r].item() - l) <= 1e-8 * abs(l), (r, lml[r].item(), l)
        np.testing.assert_allclose(grad[r].numpy(), g, rtol=1e-6, atol=1e-7 * np.abs(g).max())


def test_linear_matern_map_fit_finds_linear_variance(engine):
    from bayesianoptimizer_b200.hyperfit import fit_map
    n, d = 300, 3
    X, y = _problem(n, d, 13)
    lo = np.log(np.array([0.025] * d + [1e-2, 1e-4, 1e-4])); hi = np.log(np.array([20.0] * d + [1e2, 1.0, 1e2]))
    th0 = np.log(np.array([[0.5] * d + [1.0, 1e-2, 0.1], [1.0] * d + [0.5, 1e-2, 1.0]]))
    best, F, _, _, _ = fit_map(engine, _cuda(X), _cuda(y), "linear_matern52", th0, lo, hi, prior=None, maxiter=40)
    f0, _ = o.lml_and_grad(X, y, LIN, np.exp(th0[0, :d]), np.exp(th0[0, d]), np.exp(th0[0, d + 1]), linear_variance=np.exp(th0[0, d + 2]))
    fb, gb = o.lml_and_grad(X, y, LIN, np.exp(best[:d]), np.exp(best[d]), np.exp(best[d + 1]), linear_variance=np.exp(best[d + 2]))
    assert abs(fb - F) <= 1e-7 * max(1.0, abs(F)) and F > f0
    free = (best > lo + 1e-9) & (best < hi - 1e-9)
    assert np.abs(gb[free]).max() <= 1e-2 * max(1.0, abs(F))         # stationary in the free coordinates


@pytest.mark.parametrize("kname,kind,v", [("matern52", o.KERNEL_MATERN52, 0.0), ("linear_matern52", LIN, 0.4), ("rbf", o.KERNEL_RBF, 0.0)])
@pytest.mark.parametrize("m", [1, 8, 11])
def test_posterior_multi_shared_factorisation(engine, kname, kind, v, m):
    """m outputs, one Cholesky: means equal m separate oracle fits; the variance is the shared one."""
    n, d, N = 400, 5, 777
    X, _ = _problem(n, d, 17)
    rng = np.random.default_rng(m)
    W = rng.standard_normal((d, m))
    Y = np.sin(3 * X @ W) + 0.01 * rng.standard_normal((n, m))
    means = rng.standard_normal(m) * 0.1
    xs = rng.random((N, d))
    engine.fit(_cuda(X), _cuda(Y[:, 0]), kname, 0.5, 1.2, 1e-3, mean=means[0], linear_variance=v)
    mu, var = engine.posterior_multi(_cuda(Y), _cuda(xs), means)
    gp = o.fit(X, Y[:, 0], kind, 0.5, 1.2, 1e-3, mean=means[0], linear_variance=v)
    omu, ovar = o.posterior_multi(gp, Y, xs, means)
    assert mu.shape == (N, m)
    for t in range(m):
        assert_posterior_close(mu[:, t].cpu().numpy(), var.cpu().numpy(), omu[:, t], ovar)
    # the fitted single-output model is untouched
    m0, v0 = engine.posterior(_cuda(xs[:20]))
    assert_posterior_close(m0.cpu().numpy(), v0.cpu().numpy(), omu[:20, 0], ovar[:20])
    mu2, none = engine.posterior_multi(_cuda(Y), _cuda(xs), means, with_variance=False)
    assert none is None and torch.equal(mu2, mu)


def test_log_transformed_eight_output_model_predicts_validation_rows(engine):
    """The exact 8-output model of Bayesian6 (log + standardise -> GP -> lognormal back-transform), shared kernel."""
    import os
    from conftest import GOLDEN_DIR
    from bayesianoptimizer_b200.simulators import DEFAULT_BOUNDS
    from bayesianoptimizer_b200.transforms import LogStandardize
    z = np.load(os.path.join(GOLDEN_DIR, "csv_cache_rows.npz"))
    lo, hi = np.array(DEFAULT_BOUNDS).T
    P, Yraw = z["params"], z["outputs"]
    n = min(600, len(P) - 100)
    U = (P - lo) / (hi - lo)
    tr = LogStandardize.fit(_cuda(Yraw[:n]))
    Z = tr.forward(_cuda(Yraw[:n]))
    engine.fit(_cuda(U[:n]), Z[:, 0].contiguous(), "linear_matern52", [0.5, 0.4, 0.6, 0.8, 0.7], 1.0, 1e-2, linear_variance=0.3)
    mu, var = engine.posterior_multi(Z, _cuda(U[n:n + 100]))
    pred = tr.inverse_mean(mu, var).cpu().numpy()
    otr = o.LogStandardize.fit(Yraw[:n])
    gp = o.fit(U[:n], otr.forward(Yraw[:n])[:, 0], LIN, [0.5, 0.4, 0.6, 0.8, 0.7], 1.0, 1e-2, linear_variance=0.3)
    omu, ovar = o.posterior_multi(gp, otr.forward(Yraw[:n]), U[n:n + 100])
    np.testing.assert_allclose(pred, otr.inverse_mean(omu, ovar[:, None]), rtol=1e-7)
    assert pred.shape == (100, 8) and np.all(np.isfinite(pred))


def test_per_dimension_linear_variance_exact_gp(engine):
    """LinearKernel(ard_num_dims=d) (Bayesian7.py:162-166: raw_variance (T, 1, d)): one variance per input dimension through
    fit / posterior / sweep / append / acq_grad against the oracle; a following scalar fit is not affected."""
    X, y = _problem(260, 5, 21)
    ls = np.array([0.5, 0.4, 0.6, 0.8, 0.7])
    v = np.array([0.9, 0.05, 0.4, 1.7, 0.0])
    gp = o.fit(X, y, LIN, ls, 1.2, 1e-3, mean=0.02, linear_variance=v)
    engine.fit(_cuda(X), _cuda(y), "linear_matern52", ls, 1.2, 1e-3, mean=0.02, linear_variance=v)
    xs = np.random.default_rng(3).random((700, 5))
    mu, var = engine.posterior(_cuda(xs))
    omu, ovar = o.posterior(gp, xs)
    assert_posterior_close(mu.cpu().numpy(), var.cpu().numpy(), omu, ovar)
    val, grad = engine.acq_grad(_cuda(xs[:16]), "ucb", 0.0, 2.0)
    for i in range(16):
        ov, og = o.acquisition_with_grad(gp, xs[i], o.ACQ_UCB, 0.0, 2.0)
        assert abs(val[i].item() - ov) <= 1e-6 * abs(ov)
        np.testing.assert_allclose(grad[i].cpu().numpy(), og, rtol=2e-6, atol=1e-8 * (1 + np.abs(og).max()))
    engine.append(_cuda(xs[0]))
    gp2 = o.append_point(gp, xs[0])
    mu, var = engine.posterior(_cuda(xs[1:200]))
    omu, ovar = o.posterior(gp2, xs[1:200])
    assert_posterior_close(mu.cpu().numpy(), var.cpu().numpy(), omu, ovar)
    with pytest.raises(ValueError):
        engine.fit(_cuda(X), _cuda(y), "linear_matern52", ls, 1.2, 1e-3, linear_variance=v[:3])
    engine.fit(_cuda(X), _cuda(y), "linear_matern52", ls, 1.2, 1e-3, mean=0.02, linear_variance=0.3)      # scalar again
    gp3 = o.fit(X, y, LIN, ls, 1.2, 1e-3, mean=0.02, linear_variance=0.3)
    mu, var = engine.posterior(_cuda(xs[:100]))
    omu, ovar = o.posterior(gp3, xs[:100])
    assert_posterior_close(mu.cpu().numpy(), var.cpu().numpy(), omu, ovar)


@pytest.mark.parametrize("mode,path", [("i8x8", 8), ("i8x7", 7), ("auto", 7)])
def test_linear_matern_sliced_sweep_against_oracle(engine, mode, path):
    """The reference's explicit kernel ScaleKernel(Linear + Matern) (Bayesian6.py:470-478) on the INT8-sliced tensor path: |k*|
    is not bounded by the output scale, so every candidate carries its own power-of-two operand scale (Cauchy-Schwarz bound);
    dense outputs and top-k against the oracle at the north-star tolerances, incl. a per-dimension linear variance."""
    from bayesianoptimizer_b200 import sobol_state
    n, d, N = 700, 5, 30_000
    X, y = _problem(n, d, 13)
    ls = np.array([0.5, 0.4, 0.6, 0.8, 0.7])
    v = np.array([0.9, 0.05, 0.4, 1.7, 0.2])
    gp = o.fit(X, y, LIN, ls, 1.3, 1e-3, mean=0.05, linear_variance=v)
    engine.fit(_cuda(X), _cuda(y), "linear_matern52", ls, 1.3, 1e-3, mean=0.05, linear_variance=v)
    engine.set_sweep_mode(mode)
    try:
        se = torch.quasirandom.SobolEngine(d, scramble=True, seed=4)
        pts = o.sobol_points(se.sobolstate.numpy(), se.shift.numpy(), 0, N)
        st = sobol_state(d, 4)
        bf = float(y.max())
        for acq, ak in (("ei", o.ACQ_EI), ("ucb", o.ACQ_UCB), ("var", o.ACQ_VAR)):
            tv, ti, mu, var, av = o.sweep(gp, pts, ak, bf, 2.0, k=8)
            vals, idx, gm, gv, ga = engine.sweep(acq, bf, 2.0, sobol=st, count=N, topk=8, return_all=True)
            assert engine.last_sweep_path() == path
            assert_posterior_close(gm.cpu().numpy(), gv.cpu().numpy(), mu, var)
            if acq == "ei":
                from conftest import assert_ei_close_conditioned
                assert_ei_close_conditioned(acq, ga.cpu().numpy(), av, mu, var, bf)
            else:
                # (UCB = mu + sqrt(beta) sigma crosses zero inside a 30 000-candidate pool: relative to its two terms there)
                np.testing.assert_allclose(ga.cpu().numpy(), av, rtol=1e-6, atol=1e-6 * 1e-3 if acq == "ucb" else 0)
            got = idx.cpu().numpy()
            for r in range(8):
                assert got[r] == ti[r] or abs(av[got[r]] - tv[r]) <= 1e-6 * abs(tv[r])
    finally:
        engine.set_sweep_mode("auto")
