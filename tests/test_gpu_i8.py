"""GPU tier (-m gpu): the INT8-sliced sweep (tcgen05 kind::i8, bo_set_sweep_mode) through the C ABI against the CPU
oracle -- same tolerances as the FP64 path (BASELINE.json:north_star: mean/variance 1e-8 rel, acquisition 1e-6 rel,
identical top-k unless the oracle's own gap is below tolerance)."""
import numpy as np
import pytest

from conftest import assert_acq_close, assert_ei_close_conditioned, assert_posterior_close, synth_problem
from oracle import gp_oracle as o

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def engine():
    from bayesianoptimizer_b200 import GPEngine
    assert torch.cuda.is_available(), "the gpu tier needs a B200"
    eng = GPEngine(torch.device("cuda", 0))
    yield eng
    eng.close()


@pytest.fixture(autouse=True)
def _restore_mode(engine):
    yield
    engine.set_sweep_mode("auto")


def _kname(kind):
    return "matern52" if int(kind) == o.KERNEL_MATERN52 else "rbf"


@pytest.mark.parametrize("mode,path", [("i8x7", 7), ("i8x8", 8)])
@pytest.mark.parametrize("n,d,N,kind", [(512, 5, 4133, o.KERNEL_MATERN52), (1000, 8, 2048, o.KERNEL_MATERN52),
                                        (300, 16, 515, o.KERNEL_RBF), (2048, 10, 640, o.KERNEL_MATERN52),
                                        (129, 3, 1, o.KERNEL_RBF)])
def test_i8_sobol_sweep_against_oracle(engine, n, d, N, kind, mode, path):
    """Pinned INT8 modes on ragged pools (N not a multiple of 64, n not a multiple of the tile, a single candidate)."""
    from bayesianoptimizer_b200 import sobol_state
    X, y = synth_problem(n, d, 1, 2)
    ls = np.linspace(0.5, 0.9, d)
    gp = o.fit(X, y, kind, ls, 1.2, 1e-3, mean=0.1)
    engine.fit(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(), _kname(kind), ls, 1.2, 1e-3, mean=0.1)
    engine.set_sweep_mode(mode)
    seed, first = 3, 1000
    se = torch.quasirandom.SobolEngine(d, scramble=True, seed=seed)
    pts = o.sobol_points(se.sobolstate.numpy(), se.shift.numpy(), first, N)
    st = sobol_state(d, seed)
    bf = float(y.max())
    for acq, ak in (("ei", o.ACQ_EI), ("logei", o.ACQ_LOGEI), ("ucb", o.ACQ_UCB), ("var", o.ACQ_VAR)):
        k = min(16, N)
        tv, ti, mu, var, av = o.sweep(gp, pts, ak, bf, 2.0, k=k, first_index=first)
        vals, idx, gm, gv, ga = engine.sweep(acq, bf, 2.0, sobol=st, first_index=first, count=N, topk=k, return_all=True)
        assert engine.last_sweep_path() == path
        assert_posterior_close(gm.cpu().numpy(), gv.cpu().numpy(), mu, var)
        if acq in ("ei", "logei"):
            assert_ei_close_conditioned(acq, ga.cpu().numpy(), av, mu, var, bf)
        else:
            assert_acq_close(acq, ga.cpu().numpy(), av)
        got = idx.cpu().numpy()
        for r in range(k):
            if got[r] != ti[r]:
                tol = 1e-6 * max(abs(tv[r]), 1e-300) + (1e-6 if acq == "logei" else 0)
                assert abs(av[got[r] - first] - tv[r]) <= tol
        np.testing.assert_array_equal(vals.cpu().numpy(), ga.cpu().numpy()[got - first])


def test_i8_explicit_candidates_and_ties(engine):
    """Explicit pool: every candidate appears 8 times -> exact ties; the slice products are exact integers and the
    recombination order is fixed, so copies score bit-identically and the lowest index wins."""
    X, y = synth_problem(400, 4, 5, 6)
    gp = o.fit(X, y, o.KERNEL_MATERN52, 0.6, 1.0, 1e-3)
    engine.fit(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(), "matern52", 0.6, 1.0, 1e-3)
    engine.set_sweep_mode("i8x7")
    base = np.random.default_rng(0).random((45, 4))
    cand = np.tile(base, (8, 1))
    vals, idx, gm, gv, av = engine.sweep("ucb", 0.0, 2.0, candidates=torch.from_numpy(cand).cuda(), topk=12, return_all=True)
    assert engine.last_sweep_path() == 7
    av = av.cpu().numpy()
    assert np.array_equal(av[:45], av[45:90]) and np.array_equal(av[:45], av[315:])
    tv, ti = o.topk(av, 12)
    assert idx.cpu().tolist() == ti.tolist() and np.array_equal(vals.cpu().numpy(), tv)
    assert idx[0].item() < 45
    mu, var = o.posterior(gp, cand)
    assert_posterior_close(gm.cpu().numpy(), gv.cpu().numpy(), mu, var)


@pytest.mark.parametrize("mode", ["fp64", "i8x7", "i8x8"])
def test_pinned_mode_shard_merge_is_bit_identical(engine, mode):
    """Candidate sharding invariant for every pinned mode: a pinned mode depends on the model only, so
    shards of any size take the same path and merge to the single-sweep list bit for bit."""
    from bayesianoptimizer_b200 import sobol_state
    X, y = synth_problem(384, 6, 21, 22)
    engine.fit(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(), "matern52", 0.7, 1.0, 1e-3)
    engine.set_sweep_mode(mode)
    st = sobol_state(6, 17)
    N, k = 50_000, 8
    v1, i1 = engine.sweep("ei", float(y.max()), sobol=st, first_index=0, count=N, topk=k)
    for G in (2, 3, 8, 13):
        per = -(-N // G)
        vs, is_ = [], []
        for r in range(G):
            lo = r * per
            cnt = max(0, min(per, N - lo))
            v, i = engine.sweep("ei", float(y.max()), sobol=st, first_index=lo, count=cnt, topk=k)
            vs.append(v.cpu().numpy()); is_.append(i.cpu().numpy())
        mv, mi = o.merge_topk(vs, is_, k)
        assert mi.tolist() == i1.cpu().tolist()
        assert np.array_equal(mv, v1.cpu().numpy())


def test_auto_mode_policy_and_sharded_resolution(engine):
    """AUTO: 7 slices (one 7-bit + six 8-bit digits: 54-bit operands, 28 products) for large pools of an eligible model, FP64 for small pools and for models the sliced path does not
    cover; the sharded helper resolves on the global pool size, so a rank with a small shard follows."""
    from bayesianoptimizer_b200 import sobol_state
    from bayesianoptimizer_b200.dist import sharded_sweep
    X, y = synth_problem(600, 5, 7, 8)
    Xd, yd = torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()
    engine.set_sweep_mode("auto")
    engine.fit(Xd, yd, "matern52", 0.6, 1.0, 1e-3)
    assert engine.resolve_sweep_mode(10**6) == "i8x7" and engine.resolve_sweep_mode(10**4) == "fp64"
    engine.fit(Xd, yd, "matern52", 0.6, 1.0, 5e-6)
    assert engine.resolve_sweep_mode(10**6) == "i8x7"            # no hyper-parameter heuristic: the per-candidate guard covers it
    engine.set_sweep_mode("i8x7")
    assert engine.resolve_sweep_mode(10**6) == "i8x7" and engine.resolve_sweep_mode(10) == "i8x7"   # pinned: the model decides
    engine.set_sweep_mode("auto")
    engine.fit(Xd, yd, "linear_matern52", 0.6, 1.0, 1e-3, linear_variance=0.3)
    assert engine.resolve_sweep_mode(10**6) == "i8x7"            # per-candidate operand scale (CTA-pair kernel)
    engine.fit(Xd[:100], yd[:100], "rbf", 0.6, 1.0, 1e-3)
    assert engine.resolve_sweep_mode(10**6) == "fp64"            # one stage of rows: not worth a pipeline
    engine.fit(Xd[:300], yd[:300], "rbf", 0.6, 1.0, 1e-3)        # 384 padded rows: AUTO waits for 512, a pinned mode does not
    assert engine.resolve_sweep_mode(10**6) == "fp64"
    engine.set_sweep_mode("i8x8")
    assert engine.resolve_sweep_mode(10**6) == "i8x8"
    engine.set_sweep_mode("auto")
    engine.fit(Xd, yd, "matern52", 0.6, 1.0, 1e-3)
    st = sobol_state(5, 3)
    total, k = 60_000, 4
    v1, i1 = engine.sweep("ei", float(y.max()), sobol=st, count=total, topk=k)
    assert engine.last_sweep_path() == 7
    # world of 8: each shard alone (7 500 candidates) would fall under AUTO's pool threshold
    parts = [sharded_sweep(engine, "ei", float(y.max()), 2.0, st, total, k, rank=r, world=1) for r in range(1)]
    assert parts[0][1].tolist() == i1.tolist()
    vs, is_ = [], []
    for r in range(8):
        lo, cnt = r * 7500, 7500
        engine.set_sweep_mode(engine.resolve_sweep_mode(total))
        v, i = engine.sweep("ei", float(y.max()), sobol=st, first_index=lo, count=cnt, topk=k)
        assert engine.last_sweep_path() == 7
        engine.set_sweep_mode("auto")
        vs.append(v.cpu().numpy()); is_.append(i.cpu().numpy())
    mv, mi = o.merge_topk(vs, is_, k)
    assert mi.tolist() == i1.cpu().tolist() and np.array_equal(mv, v1.cpu().numpy())
    assert engine.sweep_mode == "auto"


def test_low_noise_model_takes_8_slices_and_meets_the_variance_bar(engine):
    """noise = 1e-4 (the reference's floor), short length scale, candidates 1e-2 .. 1e-5 away from training rows
    (sigma^2 << k**): the case 7 slices would miss (tools/ozaki_feasibility.py) -- 8 slices stay inside 1e-8."""
    n, d = 1024, 2
    X, y = synth_problem(n, d, 11, 12)
    rng = np.random.default_rng(13)
    near = np.array([np.clip(X[(j * 37) % n] + eps * rng.standard_normal(d), 0, 1)
                     for j, eps in enumerate(np.logspace(-2, -5, 2000))])
    cand = np.vstack([rng.random((20_000 - len(near), d)), near])
    gp = o.fit(X, y, o.KERNEL_MATERN52, 0.3, 1.0, 1e-4)
    engine.fit(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(), "matern52", 0.3, 1.0, 1e-4)
    engine.set_sweep_mode("i8x8")          # (AUTO picks the same at this ratio; pinned so that the test does not sit on the threshold)
    vals, idx, gm, gv, ga = engine.sweep("var", 0.0, 2.0, candidates=torch.from_numpy(cand).cuda(), topk=4, return_all=True)
    assert engine.last_sweep_path() == 8
    mu, var = o.posterior(gp, cand)
    assert var.min() < 1e-4
    assert_posterior_close(gm.cpu().numpy(), gv.cpu().numpy(), mu, var)


def test_full_size_c3_sample_i8_against_oracle(engine):
    """BASELINE config 3 shape (n_obs = 4096, d = 8): a 20 000-candidate prefix of the headline pool on the path AUTO
    picks for it (7 slices with 8-bit lower digits), dense outputs against the oracle."""
    from bayesianoptimizer_b200 import sobol_state
    n, d, N = 4096, 8, 20_000
    X, y = synth_problem(n, d, 4, 5)
    gp = o.fit(X, y, o.KERNEL_MATERN52, 0.7, 1.0, 1e-3)
    engine.fit(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(), "matern52", 0.7, 1.0, 1e-3)
    engine.set_sweep_mode("auto")
    se = torch.quasirandom.SobolEngine(d, scramble=True, seed=6)
    pts = o.sobol_points(se.sobolstate.numpy(), se.shift.numpy(), 0, N)
    st = sobol_state(d, 6)
    bf = float(y.max())
    for acq, ak in (("ei", o.ACQ_EI), ("ucb", o.ACQ_UCB)):
        tv, ti, mu, var, av = o.sweep(gp, pts, ak, bf, 2.0, k=4)
        vals, idx, gm, gv, ga = engine.sweep(acq, bf, 2.0, sobol=st, count=N, topk=4, return_all=True)
        assert engine.last_sweep_path() == 7
        assert_posterior_close(gm.cpu().numpy(), gv.cpu().numpy(), mu, var)
        assert_acq_close(acq, ga.cpu().numpy(), av)
        assert idx.cpu().tolist() == ti.tolist()


def test_release_workspace_then_sliced_sweep_is_bit_identical(engine):
    """bo_release_workspace frees the int8 operand pack and the panels (for the co-resident simulator); the next sliced
    sweep re-creates them on demand and scores bit-identically (tools/i8_release_check.py)."""
    from bayesianoptimizer_b200 import sobol_state
    X, y = synth_problem(700, 6, 31, 32)
    engine.fit(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(), "matern52", 0.6, 1.0, 1e-3)
    engine.set_sweep_mode("auto")
    st = sobol_state(6, 5)
    v1, i1 = engine.sweep("ei", float(y.max()), sobol=st, count=40_000, topk=8)
    assert engine.last_sweep_path() == 7
    engine.release_workspace()
    v2, i2 = engine.sweep("ei", float(y.max()), sobol=st, count=40_000, topk=8)
    assert engine.last_sweep_path() == 7 and torch.equal(v1, v2) and torch.equal(i1, i2)


def test_i8_peak_probe_reports_a_tensor_rate(engine):
    X, y = synth_problem(256, 3, 1, 2)
    engine.fit(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(), "rbf", 0.5, 1.0, 1e-3)
    tops = engine.i8_peak_tops(0.2)
    assert 1000.0 < tops < 6000.0, tops          # B200 nominal dense INT8: 4500 TOP/s


@pytest.mark.parametrize("variant", ["1", "0"])
@pytest.mark.parametrize("n,d,N,kind,mode", [(512, 5, 4133, o.KERNEL_MATERN52, "i8x8"), (300, 16, 515, o.KERNEL_RBF, "i8x7"),
                                             (129, 3, 1, o.KERNEL_RBF, "i8x8"), (2304, 6, 9000, o.KERNEL_MATERN52, "i8x8")])
def test_both_sliced_kernels_on_ragged_shapes(engine, monkeypatch, n, d, N, kind, mode, variant):
    """The library picks the CTA-pair kernel (cta_group::2) from 2048 padded observations up and the one-CTA kernel below;
    BO_B200_I8_PAIR forces either one, so both are held to the oracle on ragged pools, odd row-block counts (129 -> 2 row
    blocks, 300 -> 3, 2304 -> 18) and a single candidate -- and to each other's top-k."""
    from bayesianoptimizer_b200 import sobol_state
    monkeypatch.setenv("BO_B200_I8_PAIR", variant)
    X, y = synth_problem(n, d, 41, 42)
    ls = np.linspace(0.5, 0.9, d)
    gp = o.fit(X, y, kind, ls, 1.2, 1e-3, mean=0.1)
    engine.fit(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(), _kname(kind), ls, 1.2, 1e-3, mean=0.1)
    engine.set_sweep_mode(mode)
    se = torch.quasirandom.SobolEngine(d, scramble=True, seed=5)
    pts = o.sobol_points(se.sobolstate.numpy(), se.shift.numpy(), 77, N)
    st = sobol_state(d, 5)
    bf = float(y.max())
    k = min(8, N)
    tv, ti, mu, var, av = o.sweep(gp, pts, o.ACQ_UCB, bf, 2.0, k=k, first_index=77)
    vals, idx, gm, gv, ga = engine.sweep("ucb", bf, 2.0, sobol=st, first_index=77, count=N, topk=k, return_all=True)
    assert engine.last_sweep_path() == int(mode[-1])
    assert_posterior_close(gm.cpu().numpy(), gv.cpu().numpy(), mu, var)
    assert_acq_close("ucb", ga.cpu().numpy(), av)
    got = idx.cpu().numpy()
    for r in range(k):
        assert got[r] == ti[r] or abs(av[got[r] - 77] - tv[r]) <= 1e-6 * abs(tv[r])
