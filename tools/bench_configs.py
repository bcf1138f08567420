"""Secondary measurements for the BASELINE.json configs that are not the bench.py line (C2, C4, C5 and the
C1-shaped refit+suggest), written to gpurun_out/configs_report.json.  Device timings with CUDA events /
synchronised wall clock after warm-up; roofline fractions against the live-measured FP64 DMMA peak."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesianoptimizer_b200 import GPEngine, sobol_state

def synth(n, d, sx, sy):
    X = np.random.default_rng(sx).random((n, d))
    y = np.sin(3.0 * X).sum(axis=1) + 0.05 * np.random.default_rng(sy).standard_normal(n)
    return X, (y - y.mean()) / y.std(ddof=1)

def wall(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t) * 1e3)
    return float(np.median(ts))

dev = torch.device("cuda", 0)
eng = GPEngine(dev)
peak = eng.fp64_peak_tflops(True, 0.5)
rep = {"fp64_dmma_peak_tflops": peak, "hbm_gbs_measured": 6565.5}

# ---- C2: n=512, d=5, 10^6 candidates ----
X, y = synth(512, 5, 1, 2)
Xd, yd = torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev)
fit_ms = wall(lambda: eng.fit(Xd, yd, "matern52", 0.5, 1.0, 1e-3))
st = sobol_state(5, 3)
F = 512 * 512 + 512 * (3 * 5 + 12.0)
rep["C2"] = {"n": 512, "d": 5, "pool": 1_000_000, "refit_ms": fit_ms}
for mode in ("fp64", "auto"):          # the FP64 DMMA contraction, then what AUTO picks for this pool (INT8-sliced, 7 slices)
    eng.set_sweep_mode(mode)
    for _ in range(3): eng.sweep("ei", float(y.max()), sobol=st, count=1_000_000, topk=1)
    torch.cuda.synchronize(); ms = eng.last_sweep_ms()
    rep["C2"][mode] = {"path": eng.last_sweep_path(), "sweep_ms": ms, "cand_per_s": 1e6 / ms * 1e3, "tflops_alg": 1e6 * F / ms * 1e-9,
                       "ratio_to_fp64_dmma_peak": 1e6 * F / ms * 1e-9 / peak}
eng.set_sweep_mode("auto")

# ---- C1 shape: n=3000, d=5, 10^4 candidates: refit + suggest ms ----
X, y = synth(3000, 5, 11, 12)
Xd, yd = torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev)
fit_ms = wall(lambda: eng.fit(Xd, yd, "matern52", [0.5, 0.4, 0.6, 0.8, 0.7], 1.3, 1e-3))
sug_ms = wall(lambda: eng.sweep("ei", float(y.max()), sobol=st, count=10_000, topk=10))
rep["C1_shape"] = {"n": 3000, "d": 5, "pool": 10_000, "refit_ms": fit_ms, "suggest_sweep_ms": sug_ms}

# ---- C3 fit only + UCB sweep on one wave ----
X, y = synth(4096, 8, 4, 5)
Xd, yd = torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev)
fit_ms = wall(lambda: eng.fit(Xd, yd, "matern52", 0.7, 1.0, 1e-3))
st8 = sobol_state(8, 6)
N = 148 * 128 * 8
F = 4096 * 4096 + 4096 * (3 * 8 + 12.0)
out = {}
for mode in ("fp64", "auto", "i8x8"):
    eng.set_sweep_mode(mode)
    for acq in ("ei", "ucb", "logei"):
        for _ in range(2): eng.sweep(acq, float(y.max()), 2.0, sobol=st8, count=N, topk=1)
        torch.cuda.synchronize(); ms = eng.last_sweep_ms()
        out[f"{acq}/{mode}"] = {"path": eng.last_sweep_path(), "sweep_ms": ms, "cand_per_s": N / ms * 1e3,
                                "ratio_to_fp64_dmma_peak": N * F / ms * 1e-9 / peak}
eng.set_sweep_mode("auto")
rep["C3"] = {"n": 4096, "d": 8, "pool_timed": N, "refit_ms": fit_ms, "sweeps": out,
             "chol_flops": 4096 ** 3 / 3, "note": "refit = Gram + Cholesky + explicit inverse + alpha + repack"}

# ---- C4: Kriging-believer appends from n=4096 (q=16), and at n~8192 ----
def kb(n0, q, pool):
    X, y = synth(n0, 8, 7, 5)
    eng.fit(torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev), "matern52", 0.7, 1.0, 1e-3)
    pts = torch.rand(q, 8, dtype=torch.float64, device=dev)
    torch.cuda.synchronize(); t = time.perf_counter()
    for j in range(q): eng.append(pts[j])
    torch.cuda.synchronize(); app_ms = (time.perf_counter() - t) * 1e3 / q
    eng.fit(torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev), "matern52", 0.7, 1.0, 1e-3)
    torch.cuda.synchronize(); t = time.perf_counter()
    for j in range(q):
        v, i = eng.sweep("logei", float(y.max()), sobol=st8, count=pool, topk=1)
        x = eng.sobol_points(st8, i)
        eng.append(x[0])
    torch.cuda.synchronize(); batch_ms = (time.perf_counter() - t) * 1e3
    npad = (n0 + 127) // 128 * 128
    ideal_us = 2 * (npad * npad / 2 * 8) / 6565.5e9 * 1e6
    return {"n0": n0, "q": q, "pool_per_resweep": pool, "append_ms": app_ms, "append_ideal_hbm_us": ideal_us,
            "q_batch_ms": batch_ms}
rep["C4"] = [kb(4096, 16, 1_000_000), kb(8176, 16, 100_000)]

# ---- C5: batched LML + gradient, R restarts, n=2048, d=10 ----
X, y = synth(2048, 10, 8, 5)
Xd, yd = torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev)
rng = np.random.default_rng(9)
R = int(os.environ.get("C5_R", 64))
th = np.concatenate([rng.uniform(np.log(0.05), np.log(5), (R, 10)), np.zeros((R, 1)), rng.uniform(np.log(1e-4), np.log(1e-1), (R, 1))], axis=1)
eng.lml_grad_batched(Xd, yd, th)          # warm-up at the full R: the slot workspaces are sized by the first call
torch.cuda.synchronize(); t = time.perf_counter()
lml, grad, status = eng.lml_grad_batched(Xd, yd, th)
torch.cuda.synchronize(); ms = (time.perf_counter() - t) * 1e3
rep["C5"] = {"n": 2048, "d": 10, "R": R, "ms_total": ms, "ms_per_restart": ms / R, "failed": int((status != 0).sum()),
             "flops_per_restart": 2048 ** 3, "tflops": R * 2048 ** 3 / ms * 1e-9, "frac_of_fp64_peak": R * 2048 ** 3 / ms * 1e-9 / peak}
# ---- N2: SVGP pool scan, T = 8 tasks x M = 2048 inducing points, d = 5 (Bayesian7 defaults: pool 10^4) ----
from bayesianoptimizer_b200.svgp import BatchSVGPPredictor, SVGPTaskState
T, M, d5 = 8, 2048, 5
g = np.random.default_rng(21)
states = []
for _ in range(T):
    Ls = np.tril(g.standard_normal((M, M)) * 0.01) + np.diag(0.3 + 0.5 * g.random(M))
    states.append(SVGPTaskState(torch.from_numpy(g.standard_normal((M, d5))).to(dev), torch.from_numpy(g.standard_normal(M)).to(dev),
                                torch.from_numpy(Ls).to(dev), torch.full((d5,), 1.5, dtype=torch.float64), 1.0, 0.2, 0.0, 1e-3))
torch.cuda.synchronize(); t = time.perf_counter()
pred = BatchSVGPPredictor(dev, states, jitter=1e-4)
torch.cuda.synchronize(); load_ms = (time.perf_counter() - t) * 1e3
sv = {"tasks": T, "M": M, "d": d5, "load_ms_all_tasks": load_ms, "flop_per_candidate_task": 2.0 * M * M + M * (3 * d5 + 14.0)}
for pool in (10_000, 1_000_000):
    U = torch.rand(pool, d5, dtype=torch.float64, device=dev)
    ms = wall(lambda: pred.variance_score(U, 1e-3))
    sel_ms = wall(lambda: pred.select_batch(U, 500, min_variance=1e-3)) if pool == 10_000 else None
    sv[f"pool_{pool}"] = {"scan_ms": ms, "cand_per_s": pool / ms * 1e3,
                          "tflops_alg": pool * T * sv["flop_per_candidate_task"] / ms * 1e-9,
                          "frac_of_fp64_peak": pool * T * sv["flop_per_candidate_task"] / ms * 1e-9 / peak,
                          "scan_topk_fps500_ms": sel_ms}
rep["N2_svgp"] = sv
pred.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rep, open("gpurun_out/configs_report.json", "w"), indent=1)
print(json.dumps(rep, indent=1))
