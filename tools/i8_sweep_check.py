"""GPU check of the INT8-sliced sweep (BO_B200_SWEEP_IMPL=i8) against the FP64 DMMA sweep and the CPU oracle.

    python tools/i8_sweep_check.py [n] [d] [N] [slices]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianoptimizer_b200 import GPEngine, sobol_state  # noqa: E402
from oracle import gp_oracle as o  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    N = int(sys.argv[3]) if len(sys.argv) > 3 else 200_000
    S = sys.argv[4] if len(sys.argv) > 4 else "7"
    rng = np.random.default_rng(4)
    X = rng.random((n, d))
    y = np.sin(3.0 * X).sum(axis=1) + 0.05 * np.random.default_rng(5).standard_normal(n)
    y = (y - y.mean()) / y.std(ddof=1)
    eng = GPEngine(torch.device("cuda", 0))
    eng.fit(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(), "matern52", 0.7, 1.0, 1e-3)
    st = sobol_state(d, 6)
    best = float(y.max())
    res = {}
    for impl in ("fp64", "i8"):
        eng.set_sweep_mode("fp64" if impl == "fp64" else "i8x" + S)
        for rep in range(2):
            out = eng.sweep("ei", best, sobol=st, count=N, topk=8, return_all=True)
            torch.cuda.synchronize()
        res[impl] = [t.cpu().numpy() for t in out] + [eng.last_sweep_ms()]
        print(f"{impl}: sweep kernel {res[impl][-1]:.3f} ms  ({N / res[impl][-1] / 1e3:.3f} M cand/s)  top idx {res[impl][1][:4].tolist()}")
    v0, i0, m0, var0, a0, _ = res["fp64"]
    v1, i1, m1, var1, a1, _ = res["i8"]
    print("mean  max abs diff      :", np.abs(m0 - m1).max())
    rel = np.abs(var0 - var1) / var0
    print("var   max rel diff      :", rel.max(), " median", np.median(rel), " at", int(rel.argmax()), var0[rel.argmax()], var1[rel.argmax()])
    print("acq   max rel diff      :", (np.abs(a0 - a1) / np.maximum(np.abs(a0), 1e-300)).max())
    print("top-k indices identical :", i0.tolist() == i1.tolist())
    # oracle on a slice
    M = min(N, 4096)
    se = torch.quasirandom.SobolEngine(d, scramble=True, seed=6)
    pts = o.sobol_points(se.sobolstate.numpy(), se.shift.numpy(), 0, M)
    gp = o.fit(X, y, o.KERNEL_MATERN52, 0.7, 1.0, 1e-3)
    omu, ovar = o.posterior(gp, pts)
    print("vs oracle (first %d): var rel fp64 %.3e  i8 %.3e ; mean abs fp64 %.3e  i8 %.3e" % (
        M, (np.abs(var0[:M] - ovar) / ovar).max(), (np.abs(var1[:M] - ovar) / ovar).max(),
        np.abs(m0[:M] - omu).max(), np.abs(m1[:M] - omu).max()))
    eng.close()


if __name__ == "__main__":
    main()
