"""GPU tier, SURVEY 8f N2: SVGP predictive sweep (optimization/Bayesian7.py:129-195, 664-688) through the C ABI --
bo_svgp_load + bo_posterior / bo_sweep (two triangular DMMA contractions per candidate block) against the CPU oracle."""
import numpy as np
import pytest

from conftest import assert_posterior_close, synth_problem
from oracle import gp_oracle as o

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

_KN = {o.KERNEL_LINEAR_MATERN52: "linear_matern52", o.KERNEL_MATERN52: "matern52", o.KERNEL_RBF: "rbf"}


@pytest.fixture(scope="module")
def engine():
    from bayesianoptimizer_b200 import GPEngine
    eng = GPEngine(torch.device("cuda", 0))
    yield eng
    eng.close()


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _task(M, d, seed, kind=o.KERNEL_LINEAR_MATERN52, jitter=1e-6, dense_chol=0.05):
    rng = np.random.default_rng(seed)
    Z = rng.standard_normal((M, d))
    Ls = np.tril(rng.standard_normal((M, M)) * dense_chol / np.sqrt(M / 64)) + np.diag(0.3 + 0.5 * rng.random(M))
    Ls += np.triu(rng.standard_normal((M, M)), 1)              # garbage above the diagonal must be ignored
    return o.SVGPTask(Z, kind, rng.uniform(0.8, 2.0, d), 1.3, 0.2 if kind == o.KERNEL_LINEAR_MATERN52 else 0.0,
                      0.1, 2e-3, jitter, rng.standard_normal(M), Ls)


def _load(engine, t):
    engine.load_svgp(_cuda(t.Z), _cuda(t.m), _cuda(t.Ls), _KN[t.kind], t.lengthscale, t.outputscale, t.linear_variance,
                     t.mean, t.noise, t.jitter)


@pytest.mark.parametrize("M,d,N,kind", [(60, 4, 500, o.KERNEL_LINEAR_MATERN52), (128, 5, 129, o.KERNEL_LINEAR_MATERN52),
                                        (300, 5, 2000, o.KERNEL_MATERN52), (257, 3, 100, o.KERNEL_RBF),
                                        (1000, 5, 3000, o.KERNEL_LINEAR_MATERN52), (2048, 5, 20000, o.KERNEL_LINEAR_MATERN52),
                                        (640, 16, 1000, o.KERNEL_LINEAR_MATERN52)])
def test_svgp_predictive_mean_and_variance(engine, M, d, N, kind):
    t = _task(M, d, M + d, kind)
    _load(engine, t)
    xs = np.random.default_rng(7).standard_normal((N, d))
    xs[0] = t.Z[3]
    mu, var = engine.posterior(_cuda(xs))
    omu, ovar = o.svgp_predict(t, xs)
    assert_posterior_close(mu.cpu().numpy(), var.cpu().numpy(), omu, ovar)
    # the pool scan's score path: dense variance + top-k of it
    vals, idx, m2, v2, av = engine.sweep("var", candidates=_cuda(xs), topk=8, return_all=True)
    if engine.last_sweep_path() == 0:
        assert torch.equal(v2, var) and torch.equal(av, var)
    else:            # AUTO took the sliced one-pass form (M >= 512, pool >= one turn of the CTA pairs): same tolerance as above
        assert_posterior_close(m2.cpu().numpy(), v2.cpu().numpy(), omu, ovar)
        assert torch.equal(av, v2)
    tv, ti = o.topk(ovar, 8)
    assert idx.cpu().tolist() == ti.tolist()


@pytest.mark.parametrize("mode,path", [("i8x7", 7), ("i8x8", 8)])
@pytest.mark.parametrize("M,d,N,kind", [(2048, 5, 20000, o.KERNEL_LINEAR_MATERN52), (1000, 5, 5000, o.KERNEL_LINEAR_MATERN52),
                                        (300, 5, 9000, o.KERNEL_MATERN52), (640, 16, 4800, o.KERNEL_RBF),
                                        (257, 3, 100, o.KERNEL_RBF)])
def test_svgp_sliced_sweep_matches_oracle(engine, mode, path, M, d, N, kind):
    """The INT8-sliced form of the SVGP sweep: ONE pass of the CTA-pair kernel over the stacked factor
    [L^-1; Ls^T L^-1] -- var = k** + jitter - ||L^-1 k*||^2 + ||(Ls^T L^-1) k*||^2 (+ noise) -- with the per-candidate guard
    covering both terms; strict 1e-8 against the oracle's two-solve evaluation (Bayesian7.py:664-682)."""
    t = _task(M, d, M + d, kind)
    _load(engine, t)
    xs = np.random.default_rng(7).standard_normal((N, d))
    xs[:4] = t.Z[:4]                                              # on top of inducing points: the cancellation case
    engine.set_sweep_mode(mode)
    try:
        assert engine.resolve_sweep_mode(N) == mode
        vals, idx, mu, var, av = engine.sweep("var", candidates=_cuda(xs), topk=8, return_all=True)
        assert engine.last_sweep_path() == path
        flagged = engine.last_sweep_flagged()
    finally:
        engine.set_sweep_mode("auto")
    omu, ovar = o.svgp_predict(t, xs)
    assert_posterior_close(mu.cpu().numpy(), var.cpu().numpy(), omu, ovar)
    assert torch.equal(av, var)
    tv, ti = o.topk(ovar, 8)
    assert idx.cpu().tolist() == ti.tolist()
    assert flagged < N // 2                                       # the guard does not simply send the pool to the FP64 pass
    # same values as the FP64 form of the same handle, to the same tolerance
    engine.set_sweep_mode("fp64")
    try:
        _, _, mu64, var64, _ = engine.sweep("var", candidates=_cuda(xs), topk=8, return_all=True)
        assert engine.last_sweep_path() == 0
    finally:
        engine.set_sweep_mode("auto")
    assert_posterior_close(mu.cpu().numpy(), var.cpu().numpy(), mu64.cpu().numpy(), var64.cpu().numpy())


def test_svgp_auto_takes_the_sliced_path_for_the_reference_pool(engine):
    """Bayesian7's shape: M = 2048 inducing points, a 10^4-candidate pool -> AUTO resolves to the sliced form (an SVGP state has
    no row-split FP64 kernel to prefer for small pools); tiny pools stay on FP64."""
    t = _task(2048, 5, 11, jitter=1e-4)
    _load(engine, t)
    assert engine.resolve_sweep_mode(10_000) == "i8x7"
    assert engine.resolve_sweep_mode(1_000) == "fp64"
    xs = np.random.default_rng(3).standard_normal((10_000, 5))
    _, _, mu, var, _ = engine.sweep("var", candidates=_cuda(xs), topk=0, min_variance=1e-3, return_all=True)
    assert engine.last_sweep_path() == 7
    omu, ovar = o.svgp_predict(t, xs, min_variance=1e-3)
    assert_posterior_close(mu.cpu().numpy(), var.cpu().numpy(), omu, ovar)


def test_svgp_sliced_sweep_guard_rescoring_on_collapsed_posterior(engine):
    """A trained q(u) with S << I: next to inducing points the variance is a cancellation down to the jitter level, the guard
    flags those candidates and the FP64 two-pass kernel re-scores them at their pool positions (idx_map) -- values and top-k
    order are those of the oracle."""
    M, d = 600, 5
    t = _task(M, d, 21, jitter=1e-6)
    t.noise = 0.0
    t.Ls = np.tril(np.random.default_rng(2).standard_normal((M, M)) * 1e-5) + np.eye(M) * 1e-3
    _load(engine, t)
    rng = np.random.default_rng(4)
    xs = np.vstack([t.Z[:200] + 1e-4 * rng.standard_normal((200, d)), t.Z[200:300] + 3e-2 * rng.standard_normal((100, d)),
                    rng.standard_normal((9700, d))])
    engine.set_sweep_mode("i8x7")
    try:
        vals, idx, mu, var, _ = engine.sweep("var", candidates=_cuda(xs), topk=8, return_all=True)
        assert engine.last_sweep_path() == 7
        flagged = engine.last_sweep_flagged()
    finally:
        engine.set_sweep_mode("auto")
    assert 100 <= flagged < 5000, flagged
    omu, ovar = o.svgp_predict(t, xs)
    # same tolerance as the FP64 form earns on this state: 1e-8 relative with the 64-ulp floor of tests/test_gpu_i8_refdata.py
    assert_posterior_close(mu.cpu().numpy(), var.cpu().numpy(), omu, ovar, var_abs=64 * np.finfo(np.float64).eps * t.outputscale * 4)
    tv, ti = o.topk(ovar, 8)
    assert idx.cpu().tolist() == ti.tolist()


def test_svgp_float32_model_settings(engine):
    """The reference trains in float32: jitter 1e-4, min_variance 1e-3; clamp is honoured."""
    t = _task(200, 5, 3, jitter=1e-4)
    t.noise = 0.0
    t.Ls = np.eye(200) * 1e-3                                     # nearly collapsed q(u): tiny variances near Z
    _load(engine, t)
    xs = np.vstack([t.Z[:20], np.random.default_rng(1).standard_normal((200, 5))])
    mu, var = engine.posterior(_cuda(xs), min_variance=1e-3)
    omu, ovar = o.svgp_predict(t, xs, min_variance=1e-3)
    assert (ovar[:20] == 1e-3).all()
    np.testing.assert_allclose(mu.cpu().numpy(), omu, rtol=1e-8, atol=1e-8)
    assert (var.cpu().numpy()[:20] == 1e-3).all()
    np.testing.assert_allclose(var.cpu().numpy()[20:], ovar[20:], rtol=1e-7)


def test_svgp_optimal_variational_posterior_equals_exact_gp_on_device(engine):
    """Z = X with the optimal q(u): the SVGP path (two contractions) reproduces the exact path (one contraction)."""
    n, d = 300, 4
    X, y = synth_problem(n, d, 5, 6)
    ls, s2, noise = np.array([0.6, 0.5, 0.7, 0.8]), 1.2, 5e-2
    K = o.kernel_matrix(X, X, o.KERNEL_MATERN52, ls, s2)
    L = np.linalg.cholesky(K)
    Sig_u = np.linalg.inv(np.linalg.inv(K) + np.eye(n) / noise)
    mu_u = Sig_u @ y / noise
    S = np.linalg.solve(L, np.linalg.solve(L, Sig_u).T)
    xs = np.random.default_rng(2).random((1000, d))
    engine.fit(_cuda(X), _cuda(y), "matern52", ls, s2, noise)
    m_ex, v_ex = engine.posterior(_cuda(xs))
    engine.load_svgp(_cuda(X), _cuda(np.linalg.solve(L, mu_u)), _cuda(np.linalg.cholesky((S + S.T) / 2)), "matern52", ls, s2,
                     0.0, 0.0, 0.0, 0.0)
    m_sv, v_sv = engine.posterior(_cuda(xs))
    np.testing.assert_allclose(m_sv.cpu().numpy(), m_ex.cpu().numpy(), rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(v_sv.cpu().numpy(), v_ex.cpu().numpy(), rtol=1e-5, atol=1e-9)
    # and bo_fit returns the handle to the exact mode
    engine.fit(_cuda(X), _cuda(y), "matern52", ls, s2, noise)
    m2, v2 = engine.posterior(_cuda(xs))
    assert torch.equal(m2, m_ex) and torch.equal(v2, v_ex)


def test_svgp_state_refuses_exact_only_entries(engine):
    from bayesianoptimizer_b200 import BoError
    t = _task(64, 3, 1)
    _load(engine, t)
    x = _cuda(np.zeros((2, 3)))
    for call in (lambda: engine.append(x[0]), lambda: engine.refine(x, "ei", 0.0, iters=2), lambda: engine.acq_grad(x, "ei", 0.0),
                 lambda: engine.posterior_multi(_cuda(np.zeros((64, 2))), x)):
        with pytest.raises(BoError):
            call()


def test_svgp_not_positive_definite_reports_pivot(engine):
    from bayesianoptimizer_b200 import NotPositiveDefiniteError
    t = _task(50, 3, 2, kind=o.KERNEL_RBF, jitter=0.0)
    t.Z[10:26] = t.Z[4]                                           # 16 copies of one inducing point, no jitter: a pivot <= 0
    with pytest.raises(NotPositiveDefiniteError):
        _load(engine, t)
    t.jitter = 1e-4
    _load(engine, t)
    mu, var = engine.posterior(_cuda(t.Z[:5]))
    omu, ovar = o.svgp_predict(t, t.Z[:5])
    np.testing.assert_allclose(mu.cpu().numpy(), omu, rtol=1e-6, atol=1e-7)


def test_batch_svgp_pool_scan_topk_fps_matches_oracle():
    """T = 8 tasks, 10^4-candidate pool (Bayesian7 defaults): score, top-K_big and FPS picks equal the restated pipeline."""
    from bayesianoptimizer_b200.simulators import DEFAULT_BOUNDS
    from bayesianoptimizer_b200.svgp import BatchSVGPPredictor, SVGPTaskState
    T, M, d, N = 8, 256, 5, 10_000
    ot = [_task(M, d, 100 + k, jitter=1e-4) for k in range(T)]
    states = [SVGPTaskState(_cuda(t.Z), _cuda(t.m), _cuda(t.Ls), torch.from_numpy(t.lengthscale), t.outputscale, t.linear_variance,
                            t.mean, t.noise) for t in ot]
    bounds = torch.tensor(DEFAULT_BOUNDS, dtype=torch.float64).t()
    rng = np.random.default_rng(0)
    xlm, xls = rng.standard_normal((1, d)), rng.uniform(0.5, 2.0, (1, d))
    pred = BatchSVGPPredictor(torch.device("cuda", 0), states, jitter=1e-4, bounds=bounds, x_log_mean=xlm, x_log_std=xls)
    U = rng.random((N, d))
    xs = o.svgp_transform_inputs(U, bounds.numpy(), xlm, xls)
    score = o.svgp_variance_score(ot, xs, min_variance=1e-3)
    got = pred.variance_score(_cuda(U), min_variance=1e-3).cpu().numpy()
    np.testing.assert_allclose(got, score, rtol=1e-8)
    mean, var = pred.predict(_cuda(U[:300]), min_variance=1e-3)
    for k in range(T):
        omu, ovar = o.svgp_predict(ot[k], xs[:300], min_variance=1e-3)
        assert_posterior_close(mean[k].cpu().numpy(), var[k].cpu().numpy(), omu, ovar)
    pts, idx = pred.select_batch(_cuda(U), 500, fps_start=0, min_variance=1e-3)
    big = torch.topk(torch.from_numpy(score), 8000).indices.numpy()           # K_big = min(max(5000, 20*500), 8000, N)
    gap_ok = np.sort(score)[-8000] - np.sort(score)[-8001] > 1e-8 * abs(np.sort(score)[-8000])
    if gap_ok:
        ref = big[o.fps(U[big], 500, 0)]
        assert idx.cpu().tolist() == ref.tolist()
    assert pts.shape == (500, d) and len(set(idx.cpu().tolist())) == 500
    pred.close()


def test_svgp_per_dimension_linear_variance(engine):
    """The (T, 1, d) raw_variance layout of LinearKernel(ard_num_dims=d) (Bayesian7.py:162-166): predictive mean / variance of
    a task with one LinearKernel variance per input dimension, against the oracle."""
    t = _task(200, 5, 77, o.KERNEL_LINEAR_MATERN52)
    t.linear_variance = np.array([0.6, 0.02, 0.3, 1.1, 0.25])
    _load(engine, t)
    xs = np.random.default_rng(9).standard_normal((900, 5))
    mu, var = engine.posterior(_cuda(xs))
    omu, ovar = o.svgp_predict(t, xs)
    assert_posterior_close(mu.cpu().numpy(), var.cpu().numpy(), omu, ovar)


def test_svgp_sliced_state_survives_release_and_mode_changes(engine):
    """bo_release_workspace drops the sliced operand pack (re-sliced on demand) but not the model state B = Ls^T L^-1; an exact fit in
    between returns the handle to the exact mode and a later bo_svgp_load rebuilds B: the sliced values are bit-identical each time."""
    t = _task(700, 5, 33, jitter=1e-5)
    xs = _cuda(np.random.default_rng(9).standard_normal((6000, 5)))

    def sliced():
        engine.set_sweep_mode("i8x7")
        try:
            out = engine.sweep("var", candidates=xs, topk=4, return_all=True)
            assert engine.last_sweep_path() == 7
        finally:
            engine.set_sweep_mode("auto")
        return out[3].clone(), out[1].clone()

    _load(engine, t)
    v0, i0 = sliced()
    engine.release_workspace()
    v1, i1 = sliced()
    assert torch.equal(v0, v1) and torch.equal(i0, i1)
    X, y = synth_problem(300, 5, 1, 2)
    engine.fit(_cuda(X), _cuda(y), "matern52", 0.6, 1.0, 1e-2)             # exact mode in between (other n, other factor)
    engine.sweep("ei", 0.5, candidates=xs, topk=1)
    _load(engine, t)
    v2, i2 = sliced()
    assert torch.equal(v0, v2) and torch.equal(i0, i2)
    omu, ovar = o.svgp_predict(t, xs.cpu().numpy())
    np.testing.assert_allclose(v2.cpu().numpy(), ovar, rtol=1e-8)
