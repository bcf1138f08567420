"""GPU tier: K5 row append (Kriging believer), K6 acquisition gradient + refinement, K7 batched LML + gradient,
each against the CPU oracle through the C ABI."""
import numpy as np
import pytest

from conftest import assert_acq_close, assert_posterior_close, synth_problem
from oracle import gp_oracle as o

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def engine():
    from bayesianoptimizer_b200 import GPEngine
    eng = GPEngine(torch.device("cuda", 0))
    yield eng
    eng.close()


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("kind,kname", [(o.KERNEL_MATERN52, "matern52"), (o.KERNEL_RBF, "rbf")])
def test_acq_value_and_gradient_against_oracle(engine, kind, kname):
    X, y = synth_problem(333, 5, 1, 2)
    ls = [0.5, 0.4, 0.6, 0.8, 0.7]
    gp = o.fit(X, y, kind, ls, 1.3, 1e-3, mean=0.05)
    engine.fit(_cuda(X), _cuda(y), kname, ls, 1.3, 1e-3, mean=0.05)
    Xq = np.random.default_rng(3).random((21, 5))
    Xq[0] = X[7]                       # on a training point
    for acq, ak in (("ei", o.ACQ_EI), ("logei", o.ACQ_LOGEI), ("ucb", o.ACQ_UCB), ("var", o.ACQ_VAR), ("mean", o.ACQ_MEAN)):
        val, grad = engine.acq_grad(_cuda(Xq), acq, 1.0, 2.0)
        val, grad = val.cpu().numpy(), grad.cpu().numpy()
        for i in range(Xq.shape[0]):
            v, g = o.acquisition_with_grad(gp, Xq[i], ak, 1.0, 2.0)
            tol = 1e-6 * abs(v) + (1e-6 if acq == "logei" else 1e-300) + (1e-9 if acq in ("mean", "ucb") else 0)
            assert abs(val[i] - v) <= tol, (acq, i, val[i], v)
            gn = np.abs(g).max()
            assert np.abs(grad[i] - g).max() <= 2e-6 * gn + 1e-12, (acq, i, grad[i], g)


def test_refine_improves_and_matches_scipy_lbfgsb(engine):
    """bo_refine vs scipy L-BFGS-B on the oracle's LogEI from the same starts (optimize_acqf stand-in)."""
    import scipy.optimize as so
    X, y = synth_problem(256, 4, 11, 12)
    gp = o.fit(X, y, o.KERNEL_MATERN52, 0.6, 1.0, 1e-3)
    engine.fit(_cuda(X), _cuda(y), "matern52", 0.6, 1.0, 1e-3)
    bf = float(y.max())
    starts = np.random.default_rng(5).random((12, 4))
    v0, _ = engine.acq_grad(_cuda(starts), "logei", bf)
    xr, vr = engine.refine(_cuda(starts), "logei", bf, iters=150)
    xr, vr, v0 = xr.cpu().numpy(), vr.cpu().numpy(), v0.cpu().numpy()
    assert np.all(vr >= v0 - 1e-12)                       # monotone
    assert np.all((xr >= 0.0) & (xr <= 1.0))              # stays in the unit box
    # the reported value is the acquisition at the reported point
    for i in range(12):
        v, _ = o.acquisition_with_grad(gp, xr[i], o.ACQ_LOGEI, bf)
        assert abs(v - vr[i]) <= 1e-6
    ref = []
    for s in starts:
        r = so.minimize(lambda z: tuple(-np.asarray(t) for t in o.acquisition_with_grad(gp, z, o.ACQ_LOGEI, bf)), s,
                        jac=True, method="L-BFGS-B", bounds=[(0.0, 1.0)] * 4, options={"maxiter": 200})
        ref.append(-r.fun)
    ref = np.array(ref)
    # parity for a stochastic multi-start optimiser: the best refined value is at least as good
    assert vr.max() >= ref.max() - 1e-3 * max(1.0, abs(ref.max()))
    assert np.mean(vr >= ref - 1e-2 * np.maximum(1.0, np.abs(ref))) >= 0.75


@pytest.mark.parametrize("n0", [100, 126, 128, 255])
def test_believer_and_observed_append(engine, n0):
    """SURVEY App. A.6 invariants and parity with the oracle's bordered update, across the 128-row padding edge."""
    d = 3
    X, y = synth_problem(n0, d, 3, 4)
    gp = o.fit(X, y, o.KERNEL_MATERN52, 0.6, 1.0, 1e-3)
    engine.fit(_cuda(X), _cuda(y), "matern52", 0.6, 1.0, 1e-3)
    xs = np.random.default_rng(9).random((200, d))
    mu0, var0 = engine.posterior(_cuda(xs), min_variance=0.0)
    a0 = engine.state()[0].cpu().numpy()
    newpts = np.random.default_rng(10).random((5, d))
    # believer appends: alpha' = [alpha; 0], mean unchanged, variance shrinks
    for x in newpts[:3]:
        engine.append(_cuda(x))
        gp = o.append_point(gp, x)
    a1 = engine.state()[0].cpu().numpy()
    assert engine.n == n0 + 3
    np.testing.assert_allclose(a1[:n0], a0, rtol=1e-7, atol=1e-9 * np.abs(a0).max())
    assert np.abs(a1[n0:]).max() <= 1e-9 * np.abs(a0).max()
    mu1, var1 = engine.posterior(_cuda(xs), min_variance=0.0)
    np.testing.assert_allclose(mu1.cpu().numpy(), mu0.cpu().numpy(), rtol=1e-8, atol=1e-8)
    assert np.all(var1.cpu().numpy() <= var0.cpu().numpy() + 1e-12)
    omu, ovar = o.posterior(gp, xs, min_variance=0.0)
    assert_posterior_close(mu1.cpu().numpy(), var1.cpu().numpy(), omu, ovar)
    # observed appends equal a refit on the extended data
    for x, yy in zip(newpts[3:], (0.4, -0.7)):
        engine.append(_cuda(x), yy)
        gp = o.append_point(gp, x, yy)
    mu2, var2 = engine.posterior(_cuda(xs))
    omu, ovar = o.posterior(gp, xs)
    assert_posterior_close(mu2.cpu().numpy(), var2.cpu().numpy(), omu, ovar)
    ref = o.fit(gp.X, gp.y, o.KERNEL_MATERN52, 0.6, 1.0, 1e-3)
    alpha, L, Li = (t.cpu().numpy() for t in engine.state())
    np.testing.assert_allclose(np.diag(L), np.diag(ref.L), rtol=1e-8)
    assert np.abs(alpha - ref.alpha).max() <= 1e-7 * np.abs(ref.alpha).max()
    # the fused sweep sees the appended rows (packed L^-1 tiles were refreshed)
    vals, idx, m, v, a = engine.sweep("ei", float(y.max()), candidates=_cuda(xs), topk=3, return_all=True)
    tv, ti, _, _, oa = o.sweep(ref, xs, o.ACQ_EI, float(y.max()), k=3)
    assert_acq_close("ei", a.cpu().numpy(), oa)
    assert idx.cpu().tolist() == ti.tolist()


def test_append_duplicate_point_zero_noise_reports_pivot(engine):
    from bayesianoptimizer_b200 import NotPositiveDefiniteError
    X, y = synth_problem(50, 2, 5, 6)
    engine.fit(_cuda(X), _cuda(y), "matern52", 0.3, 1.0, 0.0)
    with pytest.raises(NotPositiveDefiniteError) as ei:
        for _ in range(8):
            engine.append(_cuda(X[10]))                  # the same point again and again, zero noise: lambda^2 <= 0
    assert ei.value.pivot >= 51
    # the model is still usable after the failed append
    m, v = engine.posterior(_cuda(X[:4]))
    assert np.all(np.isfinite(m.cpu().numpy()))


def test_kriging_believer_batch_matches_oracle(engine):
    """q = 4 believer batch: sweep -> append winner -> re-sweep; identical picks to the oracle loop."""
    from bayesianoptimizer_b200 import sobol_state
    X, y = synth_problem(200, 3, 21, 22)
    gp = o.fit(X, y, o.KERNEL_MATERN52, 0.5, 1.0, 1e-3)
    engine.fit(_cuda(X), _cuda(y), "matern52", 0.5, 1.0, 1e-3)
    st = sobol_state(3, 5)
    se = torch.quasirandom.SobolEngine(3, scramble=True, seed=5)
    pts = o.sobol_points(se.sobolstate.numpy(), se.shift.numpy(), 0, 3000)
    bf = float(y.max())
    for _ in range(4):
        v, i = engine.sweep("logei", bf, sobol=st, count=3000, topk=1)
        tv, ti, _, _, _ = o.sweep(gp, pts, o.ACQ_LOGEI, bf, k=1)
        assert i.item() == ti[0]
        x = engine.sobol_points(st, i)
        assert np.array_equal(x.cpu().numpy()[0], pts[ti[0]])
        engine.append(x[0])
        gp = o.append_point(gp, pts[ti[0]])


@pytest.mark.parametrize("n,d,kind,kname", [(300, 5, o.KERNEL_MATERN52, "matern52"), (700, 10, o.KERNEL_MATERN52, "matern52"),
                                            (257, 3, o.KERNEL_RBF, "rbf")])
def test_batched_lml_and_gradient_against_oracle(engine, n, d, kind, kname):
    X, y = synth_problem(n, d, 8, 9)
    rng = np.random.default_rng(9)
    R = 5
    thetas = np.concatenate([rng.uniform(np.log(0.2), np.log(3.0), (R, d)), rng.uniform(np.log(0.5), np.log(2.0), (R, 1)),
                             rng.uniform(np.log(1e-3), np.log(1e-1), (R, 1))], axis=1)
    lml, grad, status = engine.lml_grad_batched(_cuda(X), _cuda(y), thetas, kname)
    assert status.tolist() == [0] * R
    for r in range(R):
        l, g = o.lml_and_grad(X, y, kind, np.exp(thetas[r, :d]), np.exp(thetas[r, d]), np.exp(thetas[r, d + 1]))
        assert abs(lml[r].item() - l) <= 1e-8 * abs(l), (r, lml[r].item(), l)
        np.testing.assert_allclose(grad[r].numpy(), g, rtol=1e-6, atol=1e-7 * np.abs(g).max())
    # the main engine's fitted model is untouched by the restarts
    engine.fit(_cuda(X), _cuda(y), kname, 0.5, 1.0, 1e-3)
    m0, _ = engine.posterior(_cuda(X[:8]))
    engine.lml_grad_batched(_cuda(X), _cuda(y), thetas[:1], kname)
    m1, _ = engine.posterior(_cuda(X[:8]))
    assert torch.equal(m0, m1)


def test_lockstep_map_fit_on_device_matches_oracle_scipy(engine):
    """N1: MAP hyper-parameter fit driven by the batched device LML (one bo_lml_grad_batched per optimiser step)
    reaches the optimum SciPy L-BFGS-B finds on the oracle's objective."""
    import scipy.optimize as so
    from bayesianoptimizer_b200.hyperfit import fit_map, log_prior_and_grad
    n, d = 400, 5
    X, y = synth_problem(n, d, 3, 4)
    rng = np.random.default_rng(2)
    lo = np.log(np.array([0.025] * d + [1e-2, 1e-4])); hi = np.log(np.array([20.0] * d + [1e2, 1.0]))
    th0 = np.vstack([np.log([0.5] * d + [1.0, 1e-2]), rng.uniform(np.log(0.1), np.log(3), (7, d + 2))])
    th0[1:, d + 1] = np.log(1e-2)
    best, F, ths, Fs, nev = fit_map(engine, _cuda(X), _cuda(y), "matern52", th0, lo, hi, prior="gamma", maxiter=60)

    def negF(t):
        l, g = o.lml_and_grad(X, y, o.KERNEL_MATERN52, np.exp(t[:d]), np.exp(t[d]), np.exp(t[d + 1]))
        lp, lg = log_prior_and_grad(t, d, "gamma")
        return -(l + lp[0]), -(g + lg[0])
    ref = max(-so.minimize(negF, t, jac=True, method="L-BFGS-B", bounds=list(zip(lo, hi)), options={"maxiter": 200}).fun
              for t in th0[:3])
    assert F >= ref - 1e-5 * max(1.0, abs(ref)), (F, ref)
    # the reported objective is the oracle's objective at the reported point
    assert abs(-negF(best)[0] - F) <= 1e-7 * max(1.0, abs(F))


def test_lml_batch_larger_than_slot_group(engine):
    """R > 32 restarts run as several lock-step groups; results are independent of the grouping."""
    n, d = 256, 4
    X, y = synth_problem(n, d, 5, 6)
    rng = np.random.default_rng(3)
    R = 70
    th = np.concatenate([rng.uniform(np.log(0.2), np.log(2), (R, d)), rng.uniform(-0.5, 0.5, (R, 1)),
                         rng.uniform(np.log(1e-3), np.log(1e-1), (R, 1))], axis=1)
    lml, grad, st = engine.lml_grad_batched(_cuda(X), _cuda(y), th, "rbf")
    assert st.tolist() == [0] * R
    for r in (0, 31, 32, 69):
        l1, g1, _ = engine.lml_grad_batched(_cuda(X), _cuda(y), th[r:r + 1], "rbf")
        assert l1[0].item() == lml[r].item() and torch.equal(g1[0], grad[r])
        l, g = o.lml_and_grad(X, y, o.KERNEL_RBF, np.exp(th[r, :d]), np.exp(th[r, d]), np.exp(th[r, d + 1]))
        assert abs(lml[r].item() - l) <= 1e-8 * abs(l)


@pytest.mark.parametrize("N,d,m", [(8000, 5, 500), (1000, 3, 1000), (77, 16, 5), (5000, 5, 1)])
def test_device_fps_matches_reference_algorithm(engine, N, d, m):
    """bo_fps vs the greedy FPS of optimization/Bayesian7.py:82-107 (restated in the oracle), incl. duplicate points."""
    X = np.random.default_rng(N).random((N, d))
    X[N // 2] = X[3]                                   # an exact duplicate -> zero-distance tie handling
    got = engine.fps(_cuda(X), m, start=7).cpu().numpy()
    ref = o.fps(X, m, start=7)
    assert np.array_equal(got, ref)
    if m > 1:
        assert len(set(got.tolist())) == min(m, N - 1) or m == N     # distinct picks until only duplicates remain


@pytest.mark.parametrize("N,K,kind", [(10_000, 8000, "random"), (10_000, 5000, "ties"), (300, 8192, "random"), (1, 1, "random"),
                                      (1_000_000, 8192, "random"), (50_000, 4096, "allequal"), (70_000, 100, "nan"),
                                      (4097, 4096, "ties"), (20_000, 7, "signed_zero")])
def test_device_topk_scores_matches_oracle_order(engine, N, K, kind):
    """bo_topk_scores vs the oracle's (value desc, index asc, NaN last) order, incl. massive ties at the threshold."""
    rng = np.random.default_rng(N + K)
    s = rng.standard_normal(N)
    if kind == "ties":
        s = np.round(s * 4) / 4                                   # ~30 distinct values: the threshold bucket holds thousands of ties
    elif kind == "allequal":
        s[:] = 1e-3                                               # e.g. every variance clamped at min_variance
    elif kind == "nan":
        s[rng.integers(0, N, N // 2)] = np.nan
        s[:50] = -np.inf
    elif kind == "signed_zero":
        s = np.where(rng.random(N) < 0.5, 0.0, -0.0) * 1.0
        s[100] = 1.0
    vals, idx = engine.topk_scores(_cuda(s), K, first_index=123)
    tv, ti = o.topk(s, K, first_index=123)
    vals, idx = vals.cpu().numpy(), idx.cpu().numpy()
    m = len(ti)
    assert np.array_equal(idx[:m], ti)
    assert np.array_equal(vals[:m], tv)
    assert np.all(idx[m:] == -1) and np.all(np.isneginf(vals[m:]))


def test_c5_full_size_256_restarts(engine):
    """BASELINE config 5 at full size: 256 batched restarts, n_obs = 2048, d = 10 (SURVEY 8d inputs); three restarts
    against the oracle, and a restart evaluated alone equals its value inside the batch (slot / group independence)."""
    n, d, R = 2048, 10, 256
    X, y = synth_problem(n, d, 8, 5)
    rng = np.random.default_rng(9)
    th = np.concatenate([rng.uniform(np.log(0.05), np.log(5), (R, d)), np.zeros((R, 1)),
                         rng.uniform(np.log(1e-4), np.log(1e-1), (R, 1))], axis=1)
    lml, grad, st = engine.lml_grad_batched(_cuda(X), _cuda(y), th)
    assert st.tolist() == [0] * R and torch.isfinite(lml).all() and torch.isfinite(grad).all()
    for r in (0, 100, 255):
        l, g = o.lml_and_grad(X, y, o.KERNEL_MATERN52, np.exp(th[r, :d]), np.exp(th[r, d]), np.exp(th[r, d + 1]))
        assert abs(lml[r].item() - l) <= 1e-8 * abs(l), (r, lml[r].item(), l)
        np.testing.assert_allclose(grad[r].numpy(), g, rtol=1e-6, atol=1e-7 * np.abs(g).max())
    l1, g1, _ = engine.lml_grad_batched(_cuda(X), _cuda(y), th[100:101])
    assert l1[0].item() == lml[100].item() and torch.equal(g1[0], grad[100])


def test_c4_appends_across_the_4096_boundary(engine):
    """BASELINE config 4 shape: Kriging-believer / observed appends growing n past a capacity and padding boundary
    (4090 -> 4100 rows); the appended model equals an oracle refit on the extended data."""
    n0, d = 4090, 8
    X, y = synth_problem(n0 + 10, d, 7, 5)
    engine.fit(_cuda(X[:n0]), _cuda(y[:n0]), "matern52", 0.7, 1.0, 1e-3)
    for i in range(n0, n0 + 10):
        engine.append(_cuda(X[i]), float(y[i]))
    assert engine.n == n0 + 10
    ref = o.fit(X, y, o.KERNEL_MATERN52, 0.7, 1.0, 1e-3)
    xs = np.random.default_rng(3).random((300, d))
    mu, var = engine.posterior(_cuda(xs))
    omu, ovar = o.posterior(ref, xs)
    assert_posterior_close(mu.cpu().numpy(), var.cpu().numpy(), omu, ovar)
    vals, idx = engine.sweep("ei", float(y.max()), candidates=_cuda(xs), topk=4)
    tv, ti, _, _, _ = o.sweep(ref, xs, o.ACQ_EI, float(y.max()), k=4)
    assert idx.cpu().tolist() == ti.tolist()
