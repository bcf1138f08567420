"""GPU tier (-m gpu): the DEFAULT sweep contraction (AUTO -> 7 INT8 slices with 8-bit lower digits + accuracy guard) and the pinned sliced modes on
the reference's own data -- rows of results/optimization_results.csv (tests/golden/csv_*.npz: duplicated rows {12, 20},
{17, 50}, clusters, sigma^2 down to ~1e-6 of the prior variance) -- with a 20 000-candidate explicit pool that mixes
uniform points with 2 400 points 1e-2 .. 1e-5 away from training rows (conftest.refdata_pool), through the C ABI against
the CPU oracle at the north-star tolerances: mean / variance 1e-8 relative, EI / UCB STRICT 1e-6 relative, same top-8.
(optimization/Bayesian.py:98-113 scores such a pool; GPConfig.candidates_pool_size = 10^6 takes this path on every suggest.)

LogEI keeps the condition-aware tolerance (its log-space tail amplifies sigma's 1e-8 by 1 + |u| + u^2; the FP64 DMMA path
needs it on the same pool: tools/i8_refdata_check.py, profiles/r02_i8_refdata_check.log)."""
import numpy as np
import pytest

from conftest import assert_acq_close, assert_ei_close_conditioned, assert_posterior_close, load_golden, refdata_pool
from oracle import gp_oracle as o

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def engine():
    from bayesianoptimizer_b200 import GPEngine
    assert torch.cuda.is_available(), "the gpu tier needs a B200"
    eng = GPEngine(torch.device("cuda", 0))
    yield eng
    eng.set_sweep_mode("auto")
    eng.close()


_ORACLE = {}


def _problem(name, noise, near_min):
    """Golden model at the given noise + its pool + the oracle's dense answers (cached per module run)."""
    key = (name, noise, near_min)
    if key not in _ORACLE:
        g = load_golden(name)
        X, y, kind = g["X"], g["y"], int(g["kind"])
        ls, s2 = g["lengthscale"], float(g["outputscale"])
        cand = refdata_pool(X, 20_000, 2_400, near_min=near_min)
        gp = o.fit(X, y, kind, ls, s2, noise)
        mu, var = o.posterior(gp, cand)
        bf = float(y.max())
        ref = {a: o.acquisition(mu, var, k, bf, beta=2.0) for a, k in (("ei", o.ACQ_EI), ("ucb", o.ACQ_UCB), ("logei", o.ACQ_LOGEI))}
        _ORACLE[key] = (X, y, kind, ls, s2, cand, mu, var, bf, ref)
    return _ORACLE[key]


# noise 1e-3 is what the goldens were minted with; 1e-4 is the reference's floor (SingleTaskGP's noise constraint on
# standardised targets).  At the floor, candidates next to the duplicated rows reach sigma^2 = 1e-6 = gpytorch's
# min_variance clamp, where two FP64 evaluations of s2 - ||u||^2 (the oracle's and the FP64 DMMA kernel's own, mode "fp64")
# differ by 1.1e-8 relative = 1.1e-14 absolute = 50 ulp of s2 (profiles/r02_i8_refdata_check.log): summation rounding of
# the n = 3000 squares, below which no FP64 implementation can go.  Those cases carry a 64-ulp(s2) absolute floor; it
# matters only for variances within 1.4x of the clamp.
ULP64 = 64 * 2.220446049250313e-16
CASES = [("csv_n512_matern", 1e-3, 1e-5, 0.0), ("csv_n512_rbf", 1e-3, 1e-5, 0.0), ("csv_n3000_matern", 1e-3, 1e-5, 0.0),
         ("csv_n512_matern", 1e-4, 1e-5, ULP64), ("csv_n3000_matern", 1e-4, 1e-5, ULP64)]


@pytest.mark.parametrize("mode", ["auto", "i8x8", "i8x7"])
@pytest.mark.parametrize("name,noise,near_min,var_ulps", CASES)
def test_sliced_sweep_on_reference_rows_strict(engine, name, noise, near_min, var_ulps, mode):
    X, y, kind, ls, s2, cand, mu, var, bf, ref = _problem(name, noise, near_min)
    assert var.min() < 1e-4 * s2 and (var < 1e-3).sum() > 300, "the pool must reach far below the prior variance"
    engine.fit(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(), "matern52" if kind == o.KERNEL_MATERN52 else "rbf", ls, s2, noise)
    engine.set_sweep_mode(mode)
    cd = torch.from_numpy(cand).cuda()
    for acq in ("ei", "ucb", "logei"):
        vals, idx, gm, gv, ga = engine.sweep(acq, bf, 2.0, candidates=cd, topk=8, return_all=True)
        assert engine.last_sweep_path() == (8 if mode == "i8x8" else 7), "the pool must take the sliced path (AUTO: 7 wide slices)"
        flagged = engine.last_sweep_flagged()
        assert flagged > 0, "points next to training rows must reach the guard's FP64 re-score"
        assert 0 < flagged < 10_000, flagged              # ... and only those: most of the pool stays on the tensor path
        gm, gv, ga = gm.cpu().numpy(), gv.cpu().numpy(), ga.cpu().numpy()
        assert_posterior_close(gm, gv, mu, var, var_abs=var_ulps * s2)
        if acq == "logei":
            assert_ei_close_conditioned(acq, ga, ref[acq], mu, var, bf)
        else:
            assert_acq_close(acq, ga, ref[acq])              # strict 1e-6 relative
        tv, ti = o.topk(ref[acq], 8)
        got = idx.cpu().numpy()
        for r in range(8):
            if got[r] != ti[r]:                              # allowed only inside the oracle's own tolerance band
                tol = 1e-6 * max(abs(tv[r]), 1e-300) + (1e-6 if acq == "logei" else 0)
                assert abs(ref[acq][got[r]] - tv[r]) <= tol
        np.testing.assert_array_equal(vals.cpu().numpy(), ga[got])


def test_guard_flags_only_what_it_must(engine):
    """The same pool without its near-training points: nothing is flagged with 8 slices, and every value is the raw
    tensor-path value (bit-identical with the guard switched off is not observable through the ABI; the count is)."""
    X, y, kind, ls, s2, cand, mu, var, bf, ref = _problem("csv_n3000_matern", 1e-3, 1e-5)
    far = cand[var > 1e-3]
    engine.fit(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(), "matern52", ls, s2, 1e-3)
    engine.set_sweep_mode("i8x8")
    vals, idx, gm, gv, ga = engine.sweep("ei", bf, 2.0, candidates=torch.from_numpy(far).cuda(), topk=8, return_all=True)
    assert engine.last_sweep_path() == 8 and engine.last_sweep_flagged() == 0
    assert_posterior_close(gm.cpu().numpy(), gv.cpu().numpy(), mu[var > 1e-3], var[var > 1e-3])
    assert_acq_close("ei", ga.cpu().numpy(), ref["ei"][var > 1e-3])


def test_pool_on_top_of_the_data_is_rescored_almost_entirely(engine):
    """A low-noise model (1e-5) and a pool sitting 1e-6 away from its training rows, sigma^2 ~ 1e-5 everywhere: the guard
    sends (nearly) every candidate through the FP64 contraction -- each on its own merits, there is no whole-pool shortcut
    (that would make values depend on the shard layout)."""
    X, y, kind, ls, s2, cand, mu, var, bf, ref = _problem("csv_n512_matern", 1e-3, 1e-5)
    rng = np.random.default_rng(5)
    pool = np.clip(X[rng.integers(0, len(X), 20_000)] + 1e-6 * rng.standard_normal((20_000, X.shape[1])), 0, 1)
    gp = o.fit(X, y, kind, ls, s2, 1e-5)
    pm, pv = o.posterior(gp, pool)
    assert np.median(pv) < 1e-4
    engine.fit(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(), "matern52", ls, s2, 1e-5)
    engine.set_sweep_mode("i8x8")
    vals, idx, gm, gv, ga = engine.sweep("ucb", bf, 2.0, candidates=torch.from_numpy(pool).cuda(), topk=8, return_all=True)
    assert engine.last_sweep_path() == 8 and engine.last_sweep_flagged() > 15_000
    assert_posterior_close(gm.cpu().numpy(), gv.cpu().numpy(), pm, pv, var_abs=ULP64 * s2)
    tv, ti = o.topk(o.acquisition(pm, pv, o.ACQ_UCB, bf, beta=2.0), 8)
    got = idx.cpu().numpy()
    ucb = o.acquisition(pm, pv, o.ACQ_UCB, bf, beta=2.0)
    for r in range(8):
        assert got[r] == ti[r] or abs(ucb[got[r]] - tv[r]) <= 1e-6 * abs(tv[r])


def test_posterior_stays_fp64_under_auto_and_follows_a_pinned_mode(engine):
    """bo_posterior (model.posterior of Bayesian2.py:168-171) must not change numerics with N: AUTO keeps it on the FP64
    contraction however large the batch; a pinned sliced mode is an explicit opt-in."""
    X, y, kind, ls, s2, cand, mu, var, bf, ref = _problem("csv_n512_matern", 1e-3, 1e-5)
    engine.fit(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(), "matern52", ls, s2, 1e-3)
    cd = torch.from_numpy(cand).cuda()
    engine.set_sweep_mode("auto")
    m0, v0 = engine.posterior(cd)
    assert engine.last_sweep_path() == 0
    m1, v1 = engine.posterior(cd[:100])
    assert torch.equal(m0[:100], m1) and torch.equal(v0[:100], v1)
    engine.set_sweep_mode("i8x8")
    m2, v2 = engine.posterior(cd)
    assert engine.last_sweep_path() == 8
    assert_posterior_close(m2.cpu().numpy(), v2.cpu().numpy(), mu, var)
    assert_posterior_close(m0.cpu().numpy(), v0.cpu().numpy(), mu, var)
