"""GPU diagnostic behind tests/test_gpu_i8_refdata.py: every sweep contraction (FP64 DMMA, 8 and 7 INT8 slices, AUTO) on the
reference's own CSV rows (tests/golden/csv_*.npz: duplicate rows, clusters) with a 20 000-candidate explicit pool =
random points + 2 400 points 1e-2 .. 1e-5 away from training rows (incl. the duplicated rows {12, 20}, {17, 50}), against
the CPU oracle.  Prints the worst error of each quantity in units of its north-star tolerance.

    python tools/i8_refdata_check.py [case ...]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bayesianoptimizer_b200 import GPEngine  # noqa: E402
from conftest import load_golden, refdata_pool  # noqa: E402
from oracle import gp_oracle as o  # noqa: E402

cases = sys.argv[1:] or ["csv_n512_matern", "csv_n512_rbf", "csv_n3000_matern"]
eng = GPEngine(torch.device("cuda", 0))
for name in cases:
    g = load_golden(name)
    X, y, kind = g["X"], g["y"], int(g["kind"])
    ls, s2 = g["lengthscale"], float(g["outputscale"])
    for noise, near_min in ((1e-3, 1e-5), (1e-4, 1e-5), (1e-4, 1e-4)):
        cand = refdata_pool(X, 20_000, 2_400, near_min=near_min)
        t0 = time.time()
        gp = o.fit(X, y, kind, ls, s2, noise)
        mu, var = o.posterior(gp, cand)
        bf = float(y.max())
        ref = {"ei": o.acquisition(mu, var, o.ACQ_EI, bf), "ucb": o.acquisition(mu, var, o.ACQ_UCB, bf, beta=2.0),
               "logei": o.acquisition(mu, var, o.ACQ_LOGEI, bf)}
        u = (mu - bf) / np.sqrt(var)
        print(f"{name} noise={noise:g} near points down to {near_min:g}: n={len(y)} pool={len(cand)} sigma^2 in [{var.min():.2e}, {var.max():.2e}], {(var < 1e-4).sum()} below 1e-4, "
              f"u in [{u.min():.1f}, {u.max():.1f}]  (oracle {time.time() - t0:.1f} s)", flush=True)
        eng.fit(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(), "matern52" if kind == o.KERNEL_MATERN52 else "rbf", ls, s2, noise)
        cd = torch.from_numpy(cand).cuda()
        for mode in ("fp64", "i8x8", "i8x7", "auto"):
            eng.set_sweep_mode(mode)
            line = f"  {mode:5s}"
            for acq in ("ei", "ucb", "logei"):
                vals, idx, gm, gv, ga = eng.sweep(acq, bf, 2.0, candidates=cd, topk=8, return_all=True)
                gm, gv, ga = gm.cpu().numpy(), gv.cpu().numpy(), ga.cpu().numpy()
                if acq == "ei":
                    em = np.abs(gm - mu) / (1e-8 * np.abs(mu) + 1e-8)
                    ev = np.abs(gv - var) / (1e-8 * var)
                    line += f" path {eng.last_sweep_path()} flagged {eng.last_sweep_flagged()}: mean {em.max():.3f} var {ev.max():.3f} (at sigma^2 {var[ev.argmax()]:.1e})"
                if acq == "logei":
                    ea = np.abs(ga - ref[acq]) / 1e-6
                    cond = 1e-6 + 3e-8 * (1 + np.abs(u) + u * u)
                    line += f" | logei strict {ea.max():.3g} (u {u[ea.argmax()]:.1f}), conditioned {(np.abs(ga - ref[acq]) / cond).max():.3f}"
                else:
                    ea = np.abs(ga - ref[acq]) / (1e-6 * np.abs(ref[acq]) + 1e-300)
                    big = ref[acq] if acq == "ucb" else ref[acq] * (np.abs(u) < 8)
                    line += f" | {acq} strict {ea.max():.3g} (u {u[ea.argmax()]:.1f})"
                    if acq == "ei":
                        line += f", |u|<8: {ea[np.abs(u) < 8].max():.3g}"
                tv, ti = o.topk(ref[acq], 8)
                same = idx.cpu().tolist() == ti.tolist()
                line += f" top8 {'same' if same else 'DIFF'}"
            print(line, flush=True)
eng.close()
