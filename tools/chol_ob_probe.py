import sys, os, time, subprocess
if len(sys.argv) > 1:
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import numpy as np, torch
    from bayesianoptimizer_b200 import GPEngine
    eng = GPEngine(torch.device("cuda", 0))
    out = []
    for n, d in ((4096, 8), (3000, 5), (2048, 10)):
        X = np.random.default_rng(8).random((n, d)); y = np.sin(3 * X).sum(1); y = (y - y.mean()) / y.std(ddof=1)
        Xd, yd = torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda()
        ts = []
        for _ in range(5):
            torch.cuda.synchronize(); t = time.perf_counter(); eng.fit(Xd, yd, "matern52", 0.7, 1.0, 1e-3); torch.cuda.synchronize()
            ts.append((time.perf_counter() - t) * 1e3)
        row = f"n={n}: fit {min(ts):.2f} ms"
        rng = np.random.default_rng(9)
        for R in (4, 32):
            th = np.concatenate([rng.uniform(np.log(0.2), np.log(2), (R, d)), np.zeros((R, 1)), rng.uniform(np.log(1e-3), np.log(1e-1), (R, 1))], axis=1)
            eng.lml_grad_batched(Xd, yd, th)
            ts = []
            for _ in range(3):
                torch.cuda.synchronize(); t = time.perf_counter(); eng.lml_grad_batched(Xd, yd, th); torch.cuda.synchronize()
                ts.append((time.perf_counter() - t) * 1e3)
            row += f" | lml R={R}: {min(ts):.2f} ms"
        out.append(row)
    print(f"OB={os.environ.get('BO_B200_CHOL_OB')}: " + " ;; ".join(out), flush=True)
else:
    for ob in ("1", "2", "4", "8"):
        subprocess.run([sys.executable, __file__, "run"], env=dict(os.environ, BO_B200_CHOL_OB=ob))
