"""Randomised parity sweep on the GPU: random shapes / kernels / hyper-parameters / acquisitions against the CPU oracle
(fit, fused sweep incl. row-split, explicit and Sobol pools, appends, SVGP state, batched LML, large top-K).
Usage: python tools/fuzz_parity.py [cases] [seed] [sweep mode: auto | fp64 | i8x7 | i8x8]"""
import os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from bayesianoptimizer_b200 import GPEngine, sobol_state
from oracle import gp_oracle as o
from conftest import assert_posterior_close, assert_acq_close

cases, seed = int(sys.argv[1]) if len(sys.argv) > 1 else 100, int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
eng = GPEngine(torch.device("cuda", 0))
eng.set_sweep_mode(sys.argv[3] if len(sys.argv) > 3 else "auto")     # pinned INT8 modes: every eligible model takes the sliced path
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
KN = {o.KERNEL_MATERN52: "matern52", o.KERNEL_RBF: "rbf", o.KERNEL_LINEAR_MATERN52: "linear_matern52"}
AC = {"ei": o.ACQ_EI, "logei": o.ACQ_LOGEI, "ucb": o.ACQ_UCB, "var": o.ACQ_VAR, "mean": o.ACQ_MEAN}
fails = 0
t00 = time.time()
for c in range(cases):
    n = int(rng.choice([1, 2, 7, 63, 64, 65, 127, 128, 129, 200, 255, 256, 257, 500, 777, 1024, 1300]))
    d = int(rng.integers(1, 17))
    kind = int(rng.choice([0, 1, 2]))
    N = int(rng.choice([1, 5, 127, 128, 129, 1000, 3000, 20000]))
    ls = rng.uniform(0.3, 2.0, d) * np.sqrt(d / 3.0)
    s2, noise, mean = float(rng.uniform(0.3, 3)), float(10 ** rng.uniform(-4, -1)), float(rng.normal() * 0.2)
    v = float(rng.uniform(0.05, 1.0)) if kind == 2 else 0.0
    acq = str(rng.choice(list(AC)))
    desc = f"case {c}: n={n} d={d} kind={kind} N={N} acq={acq} noise={noise:.2g}"
    try:
        X = rng.random((n, d)); y = np.sin(3 * X).sum(1) + 0.1 * rng.standard_normal(n)
        if n > 1: y = (y - y.mean()) / (y.std(ddof=1) + 1e-12)
        gp = o.fit(X, y, kind, ls, s2, noise, mean, 0.0, v)
        eng.fit(cu(X), cu(y), KN[kind], ls, s2, noise, mean=mean, linear_variance=v)
        bf, beta, k = float(y.max()), float(rng.uniform(0.5, 4)), int(rng.integers(1, 9))
        if rng.random() < 0.5:
            xs = rng.random((N, d)) * 1.2 - 0.1
            out = eng.sweep(acq, bf, beta, candidates=cu(xs), topk=k, return_all=True)
        else:
            sd = int(rng.integers(0, 1000)); first = int(rng.integers(0, 5000))
            st = sobol_state(d, sd)
            se = torch.quasirandom.SobolEngine(d, scramble=True, seed=sd)
            xs = o.sobol_points(se.sobolstate.numpy(), se.shift.numpy(), first, N)
            out = eng.sweep(acq, bf, beta, sobol=st, first_index=first, count=N, topk=k, return_all=True)
            first_idx = first
        vals, idx, mu, var, av = (t.cpu().numpy() for t in out)
        fi = locals().get("first_idx", 0); first_idx = 0
        tv, ti, omu, ovar, oav = o.sweep(gp, xs, AC[acq], bf, beta, k=k, first_index=fi)
        assert_posterior_close(mu, var, omu, ovar)
        if acq == "ei": assert_acq_close(acq, av, oav)
        if acq == "ucb":                       # mu + sqrt(beta) sigma cancels near its zero crossing: relative to the terms, not the sum
            err = np.abs(av - oav) / (1e-6 * (np.abs(omu) + np.sqrt(beta * ovar)))
            assert err.max() <= 1.0, f"ucb off by {err.max():.3g}x tolerance at {err.argmax()}"
        if acq == "logei":
            # LogEI is ill-conditioned deep in the tail: d logEI / d u ~ -u, and u inherits the 1e-8 relative error of mean and sigma
            u = (omu - bf) / np.sqrt(ovar)
            tol = 1e-6 + 3e-8 * (1.0 + np.abs(u) + u * u)
            err = np.abs(av - oav) / tol
            assert err.max() <= 1.0, f"logei off by {err.max():.3g}x conditioned tolerance at {err.argmax()} (u = {u[err.argmax()]:.3g})"
        m = len(ti)
        if idx[:m].tolist() != ti.tolist():
            # accept a different pick only if the oracle's own gap is below tolerance
            srt = np.sort(oav)[::-1]
            gap = np.min(np.abs(np.diff(srt[:m + 1]))) if len(srt) > 1 else 1.0
            assert gap <= 1e-6 * max(1.0, abs(srt[0])), f"top-k differs: {idx[:m].tolist()} vs {ti.tolist()} (gap {gap:g})"
        # acquisition gradient at a few points (stationary kinds only)
        if kind != 2 and n >= 2 and rng.random() < 0.3:
            Xq = rng.random((int(rng.integers(1, 12)), d))
            val, grad = eng.acq_grad(cu(Xq), acq, bf, beta)
            val, grad = val.cpu().numpy(), grad.cpu().numpy()
            for i in range(Xq.shape[0]):
                vv, gg = o.acquisition_with_grad(gp, Xq[i], AC[acq], bf, beta)
                mq, vq = o.posterior(gp, Xq[i:i + 1])
                uq = abs((mq[0] - bf) / np.sqrt(vq[0]))
                scale = {"logei": 1.0, "ucb": abs(mq[0]) + np.sqrt(beta * vq[0]), "mean": abs(mq[0]) + 1e-3}.get(acq, abs(vv))
                cond = 1.0 + uq + uq * uq if acq in ("ei", "logei") else 1.0
                assert abs(val[i] - vv) <= 1e-6 * scale * cond + 1e-300, ("acq value", acq, val[i], vv)
                gn = np.abs(gg).max()
                if vq[0] > 1.1e-6:                      # away from the variance clamp (the clamp zeroes the variance gradient)
                    assert np.abs(grad[i] - gg).max() <= 1e-5 * gn * cond + 1e-10, ("acq grad", acq, grad[i], gg)
        # m outputs sharing the factorisation
        if n >= 4 and rng.random() < 0.2:
            m = int(rng.integers(1, 12))
            Y = np.sin(X @ rng.standard_normal((d, m))) + 0.05 * rng.standard_normal((n, m))
            means = rng.standard_normal(m) * 0.1
            q = rng.random((257, d))
            mm, vv2 = eng.posterior_multi(cu(Y), cu(q), means)
            om, ov = o.posterior_multi(gp, Y, q, means)
            for t_ in range(m):
                assert_posterior_close(mm[:, t_].cpu().numpy(), vv2.cpu().numpy(), om[:, t_], ov)
        # device farthest-point sampling on the pool
        if N >= 128 and rng.random() < 0.2:
            mfp = int(rng.integers(1, min(N, 300)))
            stt_ = int(rng.integers(0, N))
            assert np.array_equal(eng.fps(cu(xs), mfp, stt_).cpu().numpy(), o.fps(xs, mfp, stt_))
        # an append now and then (stationary and linear kinds)
        if n >= 2 and rng.random() < 0.3:
            xn = rng.random(d); yn = None if rng.random() < 0.5 else float(rng.normal())
            eng.append(cu(xn), yn); gp = o.append_point(gp, xn, yn)
            mu2, var2 = eng.posterior(cu(xs[:200])); omu2, ovar2 = o.posterior(gp, xs[:200])
            assert_posterior_close(mu2.cpu().numpy(), var2.cpu().numpy(), omu2, ovar2)
        # batched LML at a couple of random thetas
        if 8 <= n <= 600 and rng.random() < 0.3:
            R = int(rng.integers(1, 6)); p = d + 2 + (1 if kind == 2 else 0)
            th = np.concatenate([np.log(ls)[None, :] + rng.normal(0, 0.2, (R, d)), rng.normal(0, 0.3, (R, 1)), np.log(noise) + rng.normal(0, 0.3, (R, 1))] +
                                ([np.log(v) + rng.normal(0, 0.3, (R, 1))] if kind == 2 else []), axis=1)
            lml, grad, stt = eng.lml_grad_batched(cu(X), cu(y), th, KN[kind], mean)
            for r in range(R):
                l, g = o.lml_and_grad(X, y, kind, np.exp(th[r, :d]), np.exp(th[r, d]), np.exp(th[r, d + 1]), mean, np.exp(th[r, d + 2]) if kind == 2 else 0.0)
                assert stt[r] == 0 and abs(lml[r].item() - l) <= 1e-8 * max(1.0, abs(l)), (lml[r].item(), l)
                np.testing.assert_allclose(grad[r].numpy(), g, rtol=2e-6, atol=2e-7 * max(1.0, np.abs(g).max()))
        # SVGP state now and then
        if n >= 8 and rng.random() < 0.25:
            M = n
            Ls = np.tril(rng.standard_normal((M, M)) * 0.05 / np.sqrt(max(M, 64) / 64)) + np.diag(0.3 + 0.5 * rng.random(M))
            t = o.SVGPTask(X, kind, ls, s2, v, mean, noise, 1e-6 if rng.random() < 0.5 else 1e-4, rng.standard_normal(M), Ls)
            eng.load_svgp(cu(t.Z), cu(t.m), cu(t.Ls), KN[kind], ls, s2, v, mean, noise, t.jitter)
            q = rng.standard_normal((300, d)) * 0.5 + 0.5
            mu3, var3 = eng.posterior(cu(q)); omu3, ovar3 = o.svgp_predict(t, q)
            assert_posterior_close(mu3.cpu().numpy(), var3.cpu().numpy(), omu3, ovar3)
        # large top-K of the dense scores
        if N >= 1000 and rng.random() < 0.3:
            K = int(rng.choice([1, 100, 999, 4096]))
            tv2, ti2 = eng.topk_scores(cu(oav), K, 7); ov2, oi2 = o.topk(oav, K, 7)
            assert np.array_equal(ti2.cpu().numpy()[:len(oi2)], oi2)
    except Exception as e:
        fails += 1
        print("FAIL", desc, "->", repr(e)[:400], flush=True)
        if os.environ.get("FUZZ_TRACE"): traceback.print_exc()
print(f"fuzz: {cases} cases, {fails} failures, {time.time() - t00:.0f} s")
sys.exit(1 if fails else 0)
