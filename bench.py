#!/usr/bin/env python
"""Benchmark of the GP surrogate + acquisition hot path (contract: see the task's bench.py section).

    python bench.py [--config C3] [--gpus N] [--steps K] [--warmup W] [--impl reference]

--config selects a BASELINE.json configuration (C3 is the default headline: the one the metric is quoted on):
  C1  the reference's own CPU-runnable case: Bayesian.py loop on its results CSV rows, 10^4 pool     ms per BO iteration
  C2  synthetic d=5, n_obs=512, 10^6 EI candidates                       candidates/s
  C3  synthetic d=8, n_obs=4096, 10^7 EI candidates, sharded over N GPUs  candidates/s     (strong scaling: fixed pool)
  C4  Kriging-believer q=16 batches growing from n_obs=4096 (to 8192), 10^6-candidate LogEI re-sweep + one row append
      per pick                                                           ms per q=16 batch
  C5  log-marginal-likelihood + gradient over R=256 batched restarts, n_obs=2048, d=10    ms per batched evaluation
One step = one pass of the configuration's hot path over one batch of synthetic input.  `value` is timed with the inputs
resident in HBM; `e2e` goes through the host-buffer C-ABI entries (bo_fit_host / bo_sweep_host / host thetas) with the
host<->device copies inside the timed region.  N > 1 is launched by torchrun (one rank per GPU, NCCL).
--impl reference times the CPU oracle (port of the reference's exact-GP path; its botorch/gpytorch stack is not
installable offline) on the host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNIT = "candidates/s"
TOPK = 1
CONFIGS = {
    "C1": dict(n=3000, d=5, pool=10_000, seeds=(0,), golden="csv_n3000_matern.npz", q_big=1000,
               metric="GP refit+suggest ms (Bayesian.py loop, n_obs=3000 rows of the reference's results CSV, 10^4-candidate pool)",
               workload="C1: the Bayesian.py loop through the drop-in class on the first 3000 rows of the reference's "
                        "results/optimization_results.csv (tests/golden/csv_n3000_matern.npz; the Taichi simulator replaced by the cached "
                        "CSV objective): per iteration fit_gp_model (warm-started MAP hyper-parameter fit on the exact LML + final fit) -> "
                        "optimize_acquisition_function (LogEI over a 10^4 Sobol pool, 10 refined starts) -> register"),
    "C2": dict(n=512, d=5, pool=1_000_000, ls=0.5, s2=1.0, noise=1e-3, seeds=(1, 2, 3), cpu_sample=200_000,
               metric="EI candidates scored/s (n_obs=512,d=5)",
               workload="C2: synthetic d=5 n_obs=512 Matern-5/2 ARD, EI over a 10^6 in-kernel scrambled-Sobol pool"),
    "C3": dict(n=4096, d=8, pool=10_000_000, ls=0.7, s2=1.0, noise=1e-3, seeds=(4, 5, 6), cpu_sample=100_000,
               metric="EI candidates scored/s (n_obs=4096,d=8)",
               workload="C3: synthetic d=8 n_obs=4096 Matern-5/2 ARD, EI over a 10^7 in-kernel scrambled-Sobol pool sharded "
                        "contiguously over the GPUs, top-1 + one (value,index) all-gather"),
    "C4": dict(n=4096, d=8, pool=1_000_000, q=16, ls=0.7, s2=1.0, noise=1e-3, seeds=(7, 5, 6),
               metric="ms per Kriging-believer q=16 batch (n_obs from 4096, 10^6-candidate LogEI re-sweep + rank-1 append per pick)",
               workload="C4: synthetic d=8, Kriging-believer q=16 batches on a model growing from n_obs=4096: per pick one "
                        "10^6-candidate LogEI sweep (in-kernel Sobol), the winner's coordinates, one bordering append of L and L^-1"),
    "C5": dict(n=2048, d=10, R=256, seeds=(8, 5, 9),
               metric="ms per batched LML+gradient evaluation (R=256 restarts, n_obs=2048, d=10)",
               workload="C5: synthetic d=10 n_obs=2048, exact log marginal likelihood + gradient of 256 hyper-parameter restarts "
                        "in one bo_lml_grad_batched call (Gram, Cholesky, inverse, L^-T L^-1, gradient traces per restart)"),
}
# the module-level names of the headline configuration (tests and tools import them)
N_OBS, DIM, POOL = CONFIGS["C3"]["n"], CONFIGS["C3"]["d"], CONFIGS["C3"]["pool"]
METRIC = CONFIGS["C3"]["metric"]


def flops_per_candidate(n, d):
    """Algorithmic FP64 work per candidate: n^2 triangular contraction + n(3d+12) kernel row."""
    return n * n + n * (3 * d + 12.0)


def int8_ops_per_candidate(n, slices):
    """int8 operations the sliced sweep executes per candidate: every 128 x 64 stage of the lower-triangular block
    structure of the padded factor, S (S + 1) / 2 slice products, 2 ops per MAC."""
    npad = (n + 127) // 128 * 128
    return 2.0 * (slices * (slices + 1) // 2) * npad * (npad + 128) / 2


def synth_problem(cfg):
    n, d = cfg["n"], cfg["d"]
    X = np.random.default_rng(cfg["seeds"][0]).random((n, d))
    y = np.sin(3.0 * X).sum(axis=1) + 0.05 * np.random.default_rng(cfg["seeds"][1]).standard_normal(n)
    y = (y - y.mean()) / y.std(ddof=1)
    return X, y


def c5_thetas(cfg):
    rng = np.random.default_rng(cfg["seeds"][2])
    R, d = cfg["R"], cfg["d"]
    return np.concatenate([rng.uniform(np.log(0.05), np.log(5), (R, d)), np.zeros((R, 1)),
                           rng.uniform(np.log(1e-4), np.log(1e-1), (R, 1))], axis=1)


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        # samples taken under load: the upper half of the observed power draws
        order = np.argsort(pw) if pw else []
        load = [sm[i] for i in order[len(order) // 2:]] if len(order) else []
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py touches oracle/)
# ---------------------------------------------------------------------------------------------------------------------
def _blas_single():
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=1)
    except Exception:
        return None


def _blas_restore(limiter):
    if limiter is not None:
        limiter.restore_original_limits() if hasattr(limiter, "restore_original_limits") else limiter.unregister()


def cpu_sweep_run(cfg, n_cand, threads=None, acq="ei"):
    """Time the CPU oracle's sweep on a bounded sample of the workload.

    The sweep is spread over all host cores: the candidate sample is cut into one slice per core, each slice runs
    the oracle's chunked posterior + EI in its own thread (NumPy/SciPy release the GIL) with BLAS pinned to one
    thread per slice, and the per-slice top-k lists are merged -- the CPU analogue of the candidate sharding."""
    import concurrent.futures as cf
    import torch
    from oracle import gp_oracle as o
    threads = threads or (os.cpu_count() or 1)
    X, y = synth_problem(cfg)
    t0 = time.perf_counter()
    gp = o.fit(X, y, o.KERNEL_MATERN52, cfg["ls"], cfg["s2"], cfg["noise"])       # threaded LAPACK
    t_fit = time.perf_counter() - t0
    eng = torch.quasirandom.SobolEngine(cfg["d"], scramble=True, seed=cfg["seeds"][2])
    st, sh = eng.sobolstate.numpy(), eng.shift.numpy()
    best_f = float(y.max())
    ak = o.ACQ_EI if acq == "ei" else o.ACQ_LOGEI
    bounds = [(r * n_cand // threads, (r + 1) * n_cand // threads) for r in range(threads)]

    def work(lo_hi):
        lo, hi = lo_hi
        if hi <= lo:
            return np.array([-np.inf]), np.array([-1])
        pts = o.sobol_points(st, sh, lo, hi - lo)
        tv, ti, _, _, _ = o.sweep(gp, pts, ak, best_f, k=TOPK, first_index=lo)
        return tv, ti

    limiter = _blas_single()
    t0 = time.perf_counter()
    with cf.ThreadPoolExecutor(max_workers=threads) as ex:
        parts = list(ex.map(work, bounds))
    tv, ti = o.merge_topk([p[0] for p in parts], [p[1] for p in parts], TOPK)
    t_sweep = time.perf_counter() - t0
    _blas_restore(limiter)
    return {"fit_s": t_fit, "sweep_s": t_sweep, "cand_per_s": n_cand / t_sweep, "argmax": int(ti[0]), "value": float(tv[0]),
            "threads": threads, "gp": gp, "sobol": (st, sh)}


def cpu_c4_run(cfg, sample=20_000, picks=2):
    """CPU oracle, C4 on a bounded sample: `picks` Kriging-believer picks, each a `sample`-candidate LogEI sweep (all cores)
    + one bordering append; returns ms per pick at the sample size and the per-candidate / per-append split."""
    from oracle import gp_oracle as o
    r = cpu_sweep_run(cfg, sample, acq="logei")
    gp, (st, sh) = r["gp"], r["sobol"]
    sweep_s, app_s = [r["sweep_s"]], []
    for j in range(picks):
        x = o.sobol_points(st, sh, r["argmax"], 1)[0]
        t0 = time.perf_counter()
        gp = o.append_point(gp, x)
        app_s.append(time.perf_counter() - t0)
    return {"sweep_s_per_cand": float(np.mean(sweep_s)) / sample, "append_s": float(np.mean(app_s)), "threads": r["threads"],
            "sample": sample, "picks": picks}


def cpu_c5_run(cfg, restarts=4):
    """CPU oracle, C5 on a bounded sample: `restarts` of the R thetas, one LML+gradient each with all BLAS threads."""
    from oracle import gp_oracle as o
    X, y = synth_problem(cfg)
    th = c5_thetas(cfg)[:restarts]
    d = cfg["d"]
    ts = []
    for t in th:
        t0 = time.perf_counter()
        try:
            o.lml_and_grad(X, y, o.KERNEL_MATERN52, np.exp(t[:d]), float(np.exp(t[d])), float(np.exp(t[d + 1])))
        except np.linalg.LinAlgError:
            pass
        ts.append(time.perf_counter() - t0)
    return {"s_per_restart": float(np.mean(ts)), "restarts": restarts, "threads": os.cpu_count() or 1}


def run_reference(args, cfg):
    """--impl reference: the CPU path timed on the host cores with all the threads it can use, each step a bounded
    sample of the configuration's workload.  Rank 0 alone runs; the line's e2e repeats its value (nothing is copied)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    name = args.config
    base = {"impl": "reference", "metric": cfg["metric"], "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "vs_baseline": None, "dtype": "f64",
            "data": "reference CSV rows (tests/golden), cached-objective simulator" if name == "C1" else "synthetic"}
    if name in ("C2", "C3"):
        sample = cfg["cpu_sample"]
        for _ in range(1 if args.warmup > 0 else 0):
            cpu_sweep_run(cfg, 2_000)
        times = []
        for _ in range(args.steps):
            r = cpu_sweep_run(cfg, sample)
            times.append(r["sweep_s"])
        ms = 1e3 * float(np.mean(times))
        value, unit, hib, scaling = sample / (ms * 1e-3), UNIT, True, "strong"
        what = (f"{sample}-candidate prefix of the same Sobol pool per step (NumPy/SciPy FP64 oracle, one slice per core in a "
                f"thread pool, chunk 2048 like Bayesian7.py:63)")
        config = {"workload": cfg["workload"], "n_obs": cfg["n"], "d": cfg["d"], "pool": cfg["pool"], "acq": "EI"}
        cores = r["threads"]
    elif name == "C1":
        r = cpu_c1_run(cfg, steps=max(1, min(args.steps, 3)))
        ms = r["s_per_iter"] * 1e3
        value, unit, hib, scaling = ms, "ms", False, "strong"
        what = (f"{r['steps']} warm iteration(s) of the same loop at n_obs=3000, every engine call answered by the NumPy/SciPy oracle "
                f"(threaded LAPACK)")
        config = {"workload": cfg["workload"], "n_obs": cfg["n"], "d": cfg["d"], "pool": cfg["pool"], "acq": "LogEI"}
        cores = r["threads"]
    elif name == "C4":
        times = []
        for _ in range(args.steps):
            r = cpu_c4_run(cfg)
            times.append(cfg["q"] * (r["sweep_s_per_cand"] * cfg["pool"] + r["append_s"]) * 1e3)
        ms = float(np.mean(times))
        value, unit, hib, scaling = ms, "ms", False, "strong"
        what = (f"{r['picks']} picks with {r['sample']}-candidate LogEI sweeps (all cores) + bordering appends at n_obs=4096, "
                f"scaled to q=16 picks x 10^6 candidates")
        config = {"workload": cfg["workload"], "n_obs": cfg["n"], "d": cfg["d"], "pool": cfg["pool"], "q": cfg["q"], "acq": "LogEI"}
        cores = r["threads"]
    else:
        times = []
        for _ in range(args.steps):
            r = cpu_c5_run(cfg)
            times.append(r["s_per_restart"] * cfg["R"] * 1e3)
        ms = float(np.mean(times))
        value, unit, hib, scaling = ms, "ms", False, "strong"
        what = f"{r['restarts']} of the 256 restarts, one LML+gradient each with threaded LAPACK, scaled to R=256"
        config = {"workload": cfg["workload"], "n_obs": cfg["n"], "d": cfg["d"], "restarts": cfg["R"]}
        cores = r["threads"]
    line = dict(base, value=value, unit=unit, ms_per_step=ms, higher_is_better=hib, scaling=scaling, config=config,
                cpu_baseline={"value": value, "unit": unit, "cores": cores, "kind": "port",
                              "sample": what + "; the reference's botorch/gpytorch stack is not installable offline"},
                e2e={"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    _emit(line)


_JSON_FD = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on the first
    communicator when NCCL_DEBUG is set), so everything else is pointed at stderr and the line goes to the saved descriptor."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


# ---------------------------------------------------------------------------------------------------------------------
# GPU arms
# ---------------------------------------------------------------------------------------------------------------------
class Ctx:
    """Process-group plumbing shared by the configurations."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a B200: the product path has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.args = args
        self.W = max(args.warmup, 3)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce_times(self, vals):
        """max over ranks of each entry (device-timed numbers), sum of the last entry (launch count)."""
        t = self.torch.tensor(vals, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            tmax = t.clone(); self.dist.all_reduce(tmax, op=self.dist.ReduceOp.MAX)
            tsum = t.clone(); self.dist.all_reduce(tsum, op=self.dist.ReduceOp.SUM)
            out = tmax.tolist(); out[-1] = tsum[-1].item()
            return out
        return t.tolist()

    def finish(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def sweep_traffic(name, mode, count):
    """DRAM bytes per launch of the dominant kernel, from this round's `ncu --set full` capture of the SAME kernel
    instantiation (profiles/sweep_traffic.json: per configuration and contraction mode); null when there is none."""
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "sweep_traffic.json")))
        per = tj[name][mode]["dram_bytes_per_candidate"]
        return per * count, tj[name][mode].get("source")
    except Exception:
        return None, None


def sweep_roofline(eng, cfg, name, mode, count, kms, n_obs, side=None):
    """Roofline object of the fused sweep kernel that ran (`mode`): INT8 tensor pipe for the sliced kernel (denominator:
    bo_i8_peak, live), FP64 DMMA pipe for the FP64 kernel (bo_fp64_peak, live)."""
    slices = {"fp64": 0, "i8x7": 7, "i8x8": 8}[mode]
    fpc = flops_per_candidate(n_obs, cfg["d"])
    achieved = count * fpc / (kms * 1e-3) * 1e-12
    peak_tflops = eng.fp64_peak_tflops(True, 0.5)
    traffic, tsrc = sweep_traffic(name, mode, count)
    if not slices:
        return {"bound": "tensor", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops,
                "traffic": traffic, "traffic_source": tsrc, "kernel": f"sweep_kernel<{cfg['d']}> (FP64 DMMA.8x8x4 pipe)",
                "peak_source": "FP64 DMMA peak measured live by bo_fp64_peak (register-resident DMMA.8x8x4 loop); "
                               "MEASURED_PEAKS.json has no FP64 entry",
                "flop_per_candidate": fpc, "kernel_ms": kms}
    peak_tops = eng.i8_peak_tops(0.5)              # burst: a 0.5 s run of 128x256x32 MMAs on resident operands
    peak_sus = eng.i8_peak_tops(4.0)               # the same loop back to back for 4 s (power-capped clock), like the driver's bf16 figure
    ops = int8_ops_per_candidate(n_obs, slices)
    a_tops = count * ops / (kms * 1e-3) * 1e-12
    return {"bound": "tensor", "achieved": a_tops, "peak": peak_tops, "unit": "TOP/s", "frac": a_tops / peak_tops,
            "traffic": traffic, "traffic_source": tsrc,
            "kernel": (f"sweep_i8_pair_kernel<{cfg['d']}, matern52, {slices}> (tcgen05.mma.cta_group::2 kind::i8, M=256 over a CTA pair, A-collector reuse, "
                       f"INT32 accumulators in TMEM)" if (n_obs + 127) // 128 * 128 >= 2048 else
                       f"sweep_i8_kernel<{cfg['d']}, matern52, {slices}> (tcgen05.mma kind::i8 with A-collector reuse, INT32 accumulators in TMEM)"),
            "peak_source": "INT8 tensor-pipe peak measured live by bo_i8_peak (tcgen05.mma kind::i8 128x256x32 on resident operands, "
                           "0.5 s burst); MEASURED_PEAKS.json has no INT8 entry",
            "peak_sustained": peak_sus, "frac_of_sustained_peak": a_tops / peak_sus,
            "int8_ops_per_candidate": ops, "kernel_ms": kms,
            "note": "the FP64 contraction u = L^-1 k* runs as an error-free product of signed int8 slices (AUTO: one 7-bit + six 8-bit digits per "
                    "operand, 28 slice products; i8x8: eight 7-bit digits, 36 products); TMEM (512 columns) limits "
                    "the tile to 64 candidates x S accumulators; with the A operand kept in the collector across the S - s panel slices "
                    "it multiplies, a 128x64x32 kind::i8 MMA takes 37 SM cycles in isolation (floor 32; 50 without the reuse: "
                    "profiles/r02_i8_collector_probe.log); the step is long, so the sustained (power-capped) peak is the like-for-like "
                    "denominator and the burst one the strict one",
            "fp64_equivalent": {"achieved_tflops": achieved, "fp64_dmma_peak_tflops": peak_tflops,
                                "ratio_to_fp64_dmma_peak": achieved / peak_tflops, "flop_per_candidate": fpc,
                                "fp64_dmma_path": side}}


def bench_sweep(ctx, name, cfg):
    """C2 / C3: one step = one EI pass over the whole (sharded) pool + the single (value, index) exchange."""
    torch = ctx.torch
    from bayesianoptimizer_b200 import GPEngine, sobol_state
    from bayesianoptimizer_b200.dist import allgather_topk, shard_range
    args, world, rank, dev = ctx.args, ctx.world, ctx.rank, ctx.dev
    pool = int(args.pool) if args.pool else cfg["pool"]
    n, d = cfg["n"], cfg["d"]
    X, y = synth_problem(cfg)
    best_f = float(y.max())
    Xh, yh = torch.from_numpy(X).pin_memory(), torch.from_numpy(y).pin_memory()
    Xd, yd = Xh.to(dev), yh.to(dev)
    eng = GPEngine(dev)
    sob = sobol_state(d, cfg["seeds"][2])
    first, count = shard_range(pool, rank, world)
    fit = lambda Xa, ya: eng.fit(Xa, ya, "matern52", cfg["ls"], cfg["s2"], cfg["noise"])

    # ---- refit timing (device-resident inputs) ----
    fit(Xd, yd)
    torch.cuda.synchronize()
    fit_ms = []
    for _ in range(5):
        t0 = time.perf_counter()
        fit(Xd, yd)
        torch.cuda.synchronize()
        fit_ms.append((time.perf_counter() - t0) * 1e3)
    refit_ms = float(np.median(fit_ms))

    # contraction mode: resolved once on the GLOBAL pool size and pinned, so every rank's shard takes the same path
    eng.set_sweep_mode(args.sweep_mode)
    mode = eng.resolve_sweep_mode(pool)
    eng.set_sweep_mode(mode)

    def step_resident():
        v, i = eng.sweep("ei", best_f, 2.0, sobol=sob, first_index=first, count=count, topk=TOPK)
        if world > 1:
            v, i = allgather_topk(v, i, TOPK)
        return v, i

    def step_e2e():
        fit(Xh, yh)                                                            # bo_fit_host: H2D of X, y inside
        v, i = eng.sweep_host("ei", best_f, 2.0, sobol=sob, first_index=first, count=count, topk=TOPK)   # D2H inside
        if world > 1:
            v, i = allgather_topk(v.to(dev), i.to(dev), TOPK)
            v, i = v.cpu(), i.cpu()
        return v, i

    # ---- resident arm: W warm-up steps, then exactly K timed steps ----
    for _ in range(ctx.W):
        step_resident()
    ctx.barrier()
    sampler = ClockSampler(ctx.local)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    kernel_ms, flagged = [], 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    e0.record()
    for _ in range(args.steps):
        v, i = step_resident()
        kernel_ms.append(eng.last_sweep_ms())      # CUDA events on the launch stream around the fused kernel
        flagged = eng.last_sweep_flagged()
    e1.record()
    ctx.barrier()
    elapsed_ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    winner = (float(v[0].item()), int(i[0].item()))

    # ---- e2e arm: host buffers through the C-ABI host entries ----
    for _ in range(2):
        step_e2e()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ve, ie = step_e2e()
    torch.cuda.synchronize()
    e2e_local_ms = (time.perf_counter() - t0) * 1e3
    ctx.barrier()

    # ---- after the timed region: this rank's whole shard once more on the FP64 DMMA contraction -> same arg-max? ----
    side, check = None, None
    if mode != "fp64":
        eng.set_sweep_mode("fp64")
        eng.sweep("ei", best_f, 2.0, sobol=sob, first_index=first, count=min(count, 200_000), topk=TOPK)      # warm
        vf, jf = eng.sweep("ei", best_f, 2.0, sobol=sob, first_index=first, count=count, topk=TOPK)
        torch.cuda.synchronize()
        ms = eng.last_sweep_ms()
        if world > 1:
            vf, jf = allgather_topk(vf, jf, TOPK)
        side = {"value": count / (ms * 1e-3), "unit": UNIT, "pool": count, "kernel_ms": ms,
                "frac_of_fp64_dmma_peak": count * flops_per_candidate(n, d) / (ms * 1e-3) * 1e-12 / eng.fp64_peak_tflops(True, 0.3)}
        check = {"fp64_index": int(jf[0].item()), "fp64_value": float(vf[0].item()), "same_index": int(jf[0].item()) == winner[1],
                 "rel_diff": abs(float(vf[0].item()) - winner[0]) / max(abs(winner[0]), 1e-300)}
        eng.set_sweep_mode(mode)

    # the other sliced form beside the one that ran (N = 1 only): a fifth of the shard, two sweeps, the second one timed
    other = None
    if mode in ("i8x7", "i8x8") and world == 1:
        om = "i8x8" if mode == "i8x7" else "i8x7"
        eng.set_sweep_mode(om)
        if eng.resolve_sweep_mode(pool) == om:
            sub = max(count // 5, 1)
            for _ in range(2):
                eng.sweep("ei", best_f, 2.0, sobol=sob, first_index=first, count=sub, topk=TOPK)
                torch.cuda.synchronize()
            ms = eng.last_sweep_ms()
            osl = int(om[-1])
            other = {"mode": om, "value": sub / (ms * 1e-3), "unit": UNIT, "pool": sub, "kernel_ms": ms,
                     "int8_ops_per_candidate": int8_ops_per_candidate(n, osl),
                     "achieved_tops": sub * int8_ops_per_candidate(n, osl) / (ms * 1e-3) * 1e-12,
                     "note": "short run (the clock has not settled under the power cap): compare per-clock, not absolute"}
        eng.set_sweep_mode(mode)

    elapsed_ms, e2e_ms, kern_ms, launches = ctx.reduce_times([elapsed_ms, e2e_local_ms, float(np.mean(kernel_ms)), float(launches)])
    if rank == 0:
        ms_per_step = elapsed_ms / args.steps
        value = pool / (ms_per_step * 1e-3)
        e2e_value = pool / (e2e_ms / args.steps * 1e-3)
        roofline = sweep_roofline(eng, cfg, name, mode, count, float(np.mean(kernel_ms)), n, side)
        if other is not None:
            other["frac_of_int8_peak"] = other["achieved_tops"] / roofline["peak"]
            roofline["other_sliced_form"] = other
        slices = {"fp64": 0, "i8x7": 7, "i8x8": 8}[mode]
        line = {
            "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": ctx.W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["workload"], "n_obs": n, "d": d, "pool": pool, "acq": "EI",
                       "parallelism": f"candidate-shard x{world}",
                       "contraction": ("FP64 DMMA" if not slices else
                                       f"{slices} signed slices per operand ({'one 7-bit + six 8-bit digits: 54-bit operands, 28' if slices == 7 else 'eight 7-bit digits: 55-bit operands, 36'} "
                                       f"slice products) on INT8 tensor cores, exact INT32 accumulation, FP64 recombination, per-candidate accuracy guard "
                                       f"with FP64 re-score (bo_set_sweep_mode {mode}, resolved from --sweep-mode {args.sweep_mode}); "
                                       f"{flagged} candidates re-scored per step"),
                       "l2": ("inputs larger than L2: packed L^-1 + per-CTA K* panels streamed every wave" if n >= 2048 else
                              "between timed iterations every CTA rewrites its K* panels (int8 slices) and the pool index range is "
                              "re-generated in-kernel; the operands of this small model fit L2 by design")},
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(world * (n * d * 8 + n * 8 + 2052)),
                    "d2h_bytes_per_step": int(world * TOPK * 16),
                    "includes": "bo_fit_host (H2D X,y + refit) + bo_sweep_host (sweep + D2H winner) + all-gather"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "refit_ms": refit_ms, "suggest_ms": ms_per_step,
            "argmax": {"value": winner[0], "index": winner[1]},
            "argmax_check_fp64_full_pool": check,
        }
        if not args.no_cpu_baseline and world == 1:
            sample = cfg["cpu_sample"]
            r = cpu_sweep_run(cfg, sample)
            line["cpu_baseline"] = {"value": r["cand_per_s"], "unit": UNIT, "cores": r["threads"], "kind": "port",
                                    "sample": f"{sample}-candidate prefix of the same Sobol pool, NumPy/SciPy FP64 oracle, one slice "
                                              f"per core (fit {r['fit_s']:.2f} s, sweep {r['sweep_s']:.2f} s)"}
        _emit(line)
    eng.close()


def bench_c4(ctx, name, cfg):
    """C4: one step = one Kriging-believer q=16 batch on the growing model (each pick: 10^6-candidate LogEI sweep sharded
    over the ranks, winner coordinates, one bordering append on every rank's replica)."""
    torch = ctx.torch
    from bayesianoptimizer_b200 import GPEngine, sobol_state
    from bayesianoptimizer_b200.dist import allgather_topk, shard_range
    args, world, rank, dev = ctx.args, ctx.world, ctx.rank, ctx.dev
    pool = int(args.pool) if args.pool else cfg["pool"]
    n0, d, q = cfg["n"], cfg["d"], cfg["q"]
    X, y = synth_problem(cfg)
    best_f = float(y.max())
    Xh, yh = torch.from_numpy(X).pin_memory(), torch.from_numpy(y).pin_memory()
    Xd, yd = Xh.to(dev), yh.to(dev)
    eng = GPEngine(dev)
    sob = sobol_state(d, cfg["seeds"][2])
    first, count = shard_range(pool, rank, world)
    fit = lambda Xa, ya: eng.fit(Xa, ya, "matern52", cfg["ls"], cfg["s2"], cfg["noise"])
    fit(Xd, yd)
    eng.set_sweep_mode(args.sweep_mode)
    mode = eng.resolve_sweep_mode(pool)
    eng.set_sweep_mode(mode)
    kms = []

    def pick(host, solo=False):
        """solo: this rank alone sweeps the whole pool (rank 0's side measurements: no collective may be called there)."""
        f0, c0 = (0, pool) if solo else (first, count)
        if host:
            v, i = eng.sweep_host("logei", best_f, 2.0, sobol=sob, first_index=f0, count=c0, topk=TOPK)
            v, i = v.to(dev), i.to(dev)
        else:
            v, i = eng.sweep("logei", best_f, 2.0, sobol=sob, first_index=f0, count=c0, topk=TOPK)
        kms.append(eng.last_sweep_ms())
        if world > 1 and not solo:
            v, i = allgather_topk(v, i, TOPK)
        x = eng.sobol_points(sob, i[:1])
        if host:
            x = x.cpu().to(dev)                       # the pick crosses to the host (the caller logs it) and comes back
        eng.append(x[0])

    def batch(host=False, solo=False):
        for _ in range(q):
            pick(host, solo)

    for _ in range(ctx.W):
        fit(Xd, yd)
        batch()
    fit(Xd, yd)
    ctx.barrier()
    sampler = ClockSampler(ctx.local)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    kms.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    e0.record()
    for _ in range(args.steps):
        batch()                                      # the model keeps growing: n = 4096 + 16 per step
    e1.record()
    ctx.barrier()
    elapsed_ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - launches0
    n_end = eng.n
    sweep_kms = float(np.mean(kms))
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: per step a host refit at n = 4096 (bo_fit_host) + the batch through bo_sweep_host ----
    fit(Xh, yh); batch(True)
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fit(Xh, yh)
        batch(True)
    torch.cuda.synchronize()
    e2e_local_ms = (time.perf_counter() - t0) * 1e3
    ctx.barrier()

    # ---- the append alone (K5), at n = 4096 and across the 8192 boundary, against its HBM ideal ----
    hbm = measured_peaks().get("hbm_gbs", 6565.5)
    appends = []
    if rank == 0:
        for nn in (4096, 8176):
            Xa = np.random.default_rng(17).random((nn, d))
            ya = np.sin(3.0 * Xa).sum(axis=1)
            ya = (ya - ya.mean()) / ya.std(ddof=1)
            eng.fit(torch.from_numpy(Xa).to(dev), torch.from_numpy(ya).to(dev), "matern52", cfg["ls"], cfg["s2"], cfg["noise"])
            pts = torch.rand(q + 4, d, dtype=torch.float64, device=dev)
            for j in range(4):
                eng.append(pts[j])
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for j in range(q):
                eng.append(pts[4 + j])
            a1.record(); torch.cuda.synchronize()
            ms = a0.elapsed_time(a1) / q
            npad = (nn + q + 4 + 127) // 128 * 128
            bytes_alg = 2.0 * (npad * npad / 2) * 8          # L^-1 is read twice (u = L^-1 k, then L^-T u): SURVEY 8d
            appends.append({"n_obs": nn, "append_ms": ms, "algorithmic_bytes": bytes_alg, "achieved_gbs": bytes_alg / (ms * 1e-3) * 1e-9,
                            "hbm_peak_gbs": hbm, "frac_of_hbm": bytes_alg / (ms * 1e-3) * 1e-9 / hbm, "ideal_us": bytes_alg / (hbm * 1e9) * 1e6})
        # one q=16 batch that crosses n = 8192 (re-sweeps at the configured pool)
        eng.set_sweep_mode(mode)
        t0 = time.perf_counter()
        batch(solo=True)
        torch.cuda.synchronize()
        batch_8k_ms = (time.perf_counter() - t0) * 1e3

    elapsed_ms, e2e_ms, launches = ctx.reduce_times([elapsed_ms, e2e_local_ms, float(launches)])
    if rank == 0:
        ms_per_step = elapsed_ms / args.steps
        n_mid = n0 + q * args.steps // 2
        roofline = sweep_roofline(eng, cfg, name, mode, count, sweep_kms, n_mid)
        roofline["share_of_step"] = sweep_kms * q / ms_per_step
        line = {
            "metric": cfg["metric"], "value": ms_per_step, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": ctx.W,
            "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["workload"], "n_obs": n0, "n_obs_end": n_end, "d": d, "pool": pool, "q": q, "acq": "LogEI",
                       "parallelism": f"candidate-shard x{world}, appends replicated", "contraction": mode,
                       "l2": "inputs larger than L2: packed L^-1 + per-CTA K* panels streamed every wave"},
            "e2e": {"value": e2e_ms / args.steps, "unit": "ms",
                    "h2d_bytes_per_step": int(world * (n0 * d * 8 + n0 * 8 + q * (2052 + d * 8))),
                    "d2h_bytes_per_step": int(world * q * (TOPK * 16 + d * 8)),
                    "includes": "bo_fit_host at n=4096 (H2D X,y + refit) + 16 x (bo_sweep_host + D2H winner + pick D2H/H2D + bo_append)"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "per_pick_ms": ms_per_step / q, "sweep_kernel_ms": sweep_kms,
            "append": appends, "batch_crossing_8192_ms_one_gpu": batch_8k_ms,
        }
        if not args.no_cpu_baseline and world == 1:
            r = cpu_c4_run(cfg)
            v = q * (r["sweep_s_per_cand"] * pool + r["append_s"]) * 1e3
            line["cpu_baseline"] = {"value": v, "unit": "ms", "cores": r["threads"], "kind": "port",
                                    "sample": f"{r['picks']} picks with {r['sample']}-candidate LogEI sweeps (one slice per core) + bordering "
                                              f"appends (oracle.append_point: {r['append_s'] * 1e3:.0f} ms each) at n_obs=4096, scaled to "
                                              f"q=16 x 10^6 candidates"}
        _emit(line)
    eng.close()


def bench_c5(ctx, name, cfg):
    """C5: one step = one batched LML+gradient evaluation of R=256 restarts (restarts sharded over the ranks)."""
    torch = ctx.torch
    from bayesianoptimizer_b200 import GPEngine
    from bayesianoptimizer_b200.dist import sharded_lml_grad
    args, world, rank, dev = ctx.args, ctx.world, ctx.rank, ctx.dev
    n, d, R = cfg["n"], cfg["d"], cfg["R"]
    X, y = synth_problem(cfg)
    Xh, yh = torch.from_numpy(X).pin_memory(), torch.from_numpy(y).pin_memory()
    Xd, yd = Xh.to(dev), yh.to(dev)
    th = torch.from_numpy(c5_thetas(cfg))
    eng = GPEngine(dev)
    peak = eng.fp64_peak_tflops(True, 0.5)

    def step(host=False):
        Xa, ya = (Xh.to(dev, non_blocking=True), yh.to(dev, non_blocking=True)) if host else (Xd, yd)
        return sharded_lml_grad(eng, Xa, ya, th, "matern52", 0.0, rank, world)     # host thetas in, host lml/grad/status out

    for _ in range(ctx.W):
        step()
    ctx.barrier()
    sampler = ClockSampler(ctx.local)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    e0.record()
    for _ in range(args.steps):
        lml, grad, status = step()
    e1.record()
    ctx.barrier()
    elapsed_ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    step(True)
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(True)
    torch.cuda.synchronize()
    e2e_local_ms = (time.perf_counter() - t0) * 1e3
    ctx.barrier()
    elapsed_ms, e2e_ms, launches = ctx.reduce_times([elapsed_ms, e2e_local_ms, float(launches)])
    if rank == 0:
        ms_per_step = elapsed_ms / args.steps
        flop = float(R) * n ** 3                  # per restart: Cholesky n^3/3 + inverse n^3/3 + L^-T L^-1 n^3/3
        achieved = flop / (ms_per_step * 1e-3) * 1e-12
        p = d + 2
        line = {
            "metric": cfg["metric"], "value": ms_per_step, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": ctx.W,
            "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["workload"], "n_obs": n, "d": d, "restarts": R, "parallelism": f"restart-shard x{world}",
                       "l2": "inputs larger than L2: 256 restarts x (K, L^-1, workspace) of 32 MB each are streamed through HBM every step"},
            "e2e": {"value": e2e_ms / args.steps, "unit": "ms", "h2d_bytes_per_step": int(world * (n * d * 8 + n * 8) + R * p * 8),
                    "d2h_bytes_per_step": int(R * (p + 2) * 8),
                    "includes": "H2D of X, y from pinned host memory + bo_lml_grad_batched (host thetas in, host lml/grad/status out)"},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                         "kernel": "dgemm_grouped_kernel<64,64> (FP64 DMMA.8x8x4 trailing updates / inverse / L^-T L^-1 of the lock-step restarts)",
                         "peak_source": "FP64 DMMA peak measured live by bo_fp64_peak; MEASURED_PEAKS.json has no FP64 entry",
                         "flop_per_restart": float(n) ** 3, "note": "whole-step rate (all kernels of the evaluation), not one launch"},
            "ms_per_restart": ms_per_step / R, "failed_restarts": int((status != 0).sum()),
        }
        if not args.no_cpu_baseline and world == 1:
            r = cpu_c5_run(cfg)
            line["cpu_baseline"] = {"value": r["s_per_restart"] * R * 1e3, "unit": "ms", "cores": r["threads"], "kind": "port",
                                    "sample": f"{r['restarts']} of the 256 restarts (oracle.lml_and_grad, threaded LAPACK: "
                                              f"{r['s_per_restart']:.2f} s each), scaled to R=256"}
        _emit(line)
    eng.close()


def c1_problem(cfg):
    """Rows of the reference's results CSV as committed under tests/golden (normalised X, standardised objective): returned as
    physical parameters + 8 displacement columns whose mean is the objective (Bayesian.py:140), for CachedCSVSimulator."""
    from bayesianoptimizer_b200.simulators import DEFAULT_BOUNDS
    z = np.load(os.path.join(ROOT, "tests", "golden", cfg["golden"]))
    lo, hi = np.asarray(DEFAULT_BOUNDS, dtype=np.float64).T
    n = cfg["n"]
    P = z["X"][:n] * (hi - lo) + lo
    obj = z["y"][:n] * float(z["y_std"]) + float(z["y_mean"])
    Y8 = np.repeat(obj[:, None], 8, axis=1)
    hyper = (np.asarray(z["lengthscale"], dtype=np.float64), float(z["outputscale"]), float(z["noise"]), 0.0)
    return P, Y8, hyper


def c1_optimizer(cfg, out_dir, engine_factory=None, pool=None, **cfg_kw):
    from bayesianoptimizer_b200.optimizer import BayesianOptimizer, GPConfig
    from bayesianoptimizer_b200.simulators import DEFAULT_BOUNDS, CachedCSVSimulator
    P, Y8, hyper = c1_problem(cfg)
    gpc = GPConfig(seed=cfg["seeds"][0], candidates_pool_size=pool or cfg["pool"], **cfg_kw)
    kw = {} if engine_factory is None else {"engine_factory": engine_factory}
    opt = BayesianOptimizer(CachedCSVSimulator(P, Y8), DEFAULT_BOUNDS, out_dir, 0, 1, 1, gp_config=gpc, **kw)
    for i in range(len(P)):
        opt._append_observation(P[i], Y8[i], write=False)
    return opt, hyper


def c1_iteration(opt):
    """One pass of the reference's loop body (Bayesian.py:150-172): refit, suggest, evaluate + record."""
    gp = opt.fit_gp_model()
    batch = opt.optimize_acquisition_function(gp)
    for x in batch:
        opt.register(x)
    return batch


def cpu_c1_run(cfg, steps=1):
    """CPU leg of C1: the SAME host loop (drop-in class) with every engine call answered by the NumPy/SciPy oracle
    (tests/oracle_engine.py), warm-started from the fixture's hyper-parameters like iteration k > 1 of the reference's loop."""
    import contextlib
    import io
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_engine import OracleEngine
    with tempfile.TemporaryDirectory() as tmp:
        opt, hyper = c1_optimizer(cfg, tmp, engine_factory=OracleEngine)
        opt._hyper = hyper
        opt._hyper_fits = 1
        ts = []
        for _ in range(steps):
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                c1_iteration(opt)
            ts.append(time.perf_counter() - t0)
        opt.close()
    return {"s_per_iter": float(np.mean(ts)), "steps": steps, "threads": os.cpu_count() or 1}


def bench_c1(ctx, name, cfg):
    """C1: one step = one iteration of the reference's BO loop through the drop-in class (host observations in, host
    suggestion out -- the public call IS the host-buffer path, so `value` and `e2e` time the same calls, the first on the
    device timeline, the second by wall clock)."""
    import contextlib
    import io
    import tempfile
    torch = ctx.torch
    args, world, rank, dev = ctx.args, ctx.world, ctx.rank, ctx.dev
    n, d = cfg["n"], cfg["d"]
    tmp = tempfile.mkdtemp(prefix="bo_b200_c1_")
    quiet = lambda: contextlib.redirect_stdout(io.StringIO())
    opt, _ = c1_optimizer(cfg, tmp)
    eng = opt._engine_get()
    peak = eng.fp64_peak_tflops(True, 0.5)
    # count the LML+gradient evaluations (K7) and their device time: the dominant kernel group of an iteration
    stats = {"calls": 0, "restarts": 0, "ms": 0.0}
    inner = eng.lml_grad_batched

    def counted(X, y, thetas, *a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = inner(X, y, thetas, *a, **k)
        e1.record(); e1.synchronize()
        stats["calls"] += 1; stats["restarts"] += int(np.asarray(thetas).reshape(-1, np.asarray(thetas).shape[-1]).shape[0])
        stats["ms"] += e0.elapsed_time(e1)
        return out
    eng.lml_grad_batched = counted

    def timed(fn):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        with quiet():
            r = fn()
        torch.cuda.synchronize()
        return r, (time.perf_counter() - t0) * 1e3

    _, cold_first_ms = timed(opt.fit_gp_model)          # very first refit: + one-off allocation of the 16-restart LML workspace (3.4 GB)
    for k in stats:
        stats[k] = 0
    opt._hyper, opt._hyper_fits = None, 0               # forget the fit: the same cold 16-restart search again, workspaces in place
    _, cold_ms = timed(opt.fit_gp_model)                # 16 screened restarts, the best 4 refined in lock step
    cold = dict(stats)
    for _ in range(ctx.W):
        timed(lambda: c1_iteration(opt))
    ctx.barrier()
    sampler = ClockSampler(ctx.local)
    if rank == 0:
        sampler.start()
    for k in stats:
        stats[k] = 0
    launches0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    e0.record()
    with quiet():
        for _ in range(args.steps):
            c1_iteration(opt)
    e1.record()
    ctx.barrier()
    elapsed_ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - launches0
    warm = dict(stats)
    clocks = sampler.stop() if rank == 0 else None
    ctx.barrier()
    t0 = time.perf_counter()
    with quiet():
        for _ in range(args.steps):
            c1_iteration(opt)
    torch.cuda.synchronize()
    e2e_local_ms = (time.perf_counter() - t0) * 1e3
    ctx.barrier()
    # phases of one warm iteration + a q = 1000 suggestion (the reference's batch_size, main.py:15: one sweep -> top-K -> FPS)
    phases = {}
    gp, phases["fit_gp_model_ms"] = timed(opt.fit_gp_model)
    _, phases["suggest_q1_ms"] = timed(lambda: opt.suggest(1, gp))
    timed(lambda: opt.suggest(cfg["q_big"], gp))
    xs, phases[f"suggest_q{cfg['q_big']}_ms"] = timed(lambda: opt.suggest(cfg["q_big"], gp))
    phases[f"suggest_q{cfg['q_big']}_distinct"] = int(torch.unique(xs, dim=0).shape[0])
    _, phases["suggest_q16_kriging_believer_ms"] = timed(lambda: opt.suggest(16, gp))
    y_std, _, _, _ = opt._model_targets()
    Xd = opt.train_X.to(dev)
    _, phases["plain_fit_ms"] = timed(lambda: eng.fit(Xd, y_std.to(dev), "matern52", *opt._hyper[:3]))
    n_end = int(opt.train_X.shape[0])
    opt.close()
    elapsed_ms, e2e_ms, launches = ctx.reduce_times([elapsed_ms, e2e_local_ms, float(launches)])
    if rank == 0:
        ms_per_step = elapsed_ms / args.steps
        evals = warm["restarts"] / args.steps
        flop = float(n) ** 3                              # per LML+gradient evaluation
        achieved = warm["restarts"] * flop / (max(warm["ms"], 1e-9) * 1e-3) * 1e-12
        line = {
            "metric": cfg["metric"], "value": ms_per_step, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": ctx.W,
            "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "reference CSV rows (tests/golden), cached-objective simulator",
            "config": {"workload": cfg["workload"], "n_obs": n, "n_obs_end": n_end, "d": d, "pool": cfg["pool"], "acq": "LogEI",
                       "parallelism": f"replicas x{world}: pool and hyper-fit restarts sharded, observations replicated",
                       "l2": "inputs larger than L2: K, L and L^-1 of n=3000 (72 MB each) are rewritten by every LML evaluation"},
            "e2e": {"value": e2e_ms / args.steps, "unit": "ms",
                    "h2d_bytes_per_step": int((n * d * 8 + n * 8) * 2 + evals * (d + 2) * 8),
                    "d2h_bytes_per_step": int(evals * (d + 4) * 8 + 10 * (d + 1) * 8 * 2),
                    "includes": "the public calls of the reference's loop on host observations: fit_gp_model (H2D X, y; host thetas in, "
                                "host lml/grad out per L-BFGS trial) + optimize_acquisition_function (D2H suggestion) + register (CSV row)"},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                         "kernel": "bo_lml_grad_batched at R=1 (chol_panel_kernel chain + dgemm_grouped_kernel<64,64> FP64 DMMA updates)",
                         "peak_source": "FP64 DMMA peak measured live by bo_fp64_peak; MEASURED_PEAKS.json has no FP64 entry",
                         "flop_per_evaluation": flop, "evaluations_per_step": evals,
                         "share_of_step": warm["ms"] / max(elapsed_ms, 1e-9),
                         "note": "rate over the LML+gradient evaluations of the timed steps (CUDA events around each call)"},
            "hyperfit": {"first_fit_gp_model_ms_with_workspace_allocation": cold_first_ms,
                         "cold_fit_gp_model_ms": cold_ms, "cold_lml_calls": cold["calls"], "cold_lml_restart_evaluations": cold["restarts"],
                         "cold_lml_device_ms": cold["ms"], "warm_lml_calls_per_step": warm["calls"] / args.steps,
                         "warm_lml_device_ms_per_step": warm["ms"] / args.steps,
                         "host_share_of_cold_fit": 1.0 - cold["ms"] / max(cold_ms, 1e-9),
                         "note": "theta, the L-BFGS history and the line search live on the host (hyperfit.py); host_share is the part of "
                                 "the cold fit NOT spent inside bo_lml_grad_batched -- what a device-resident optimiser could remove"},
            "phases": phases,
        }
        if not args.no_cpu_baseline and world == 1:
            r = cpu_c1_run(cfg)
            line["cpu_baseline"] = {"value": r["s_per_iter"] * 1e3, "unit": "ms", "cores": r["threads"], "kind": "port",
                                    "sample": f"{r['steps']} warm iteration of the same loop at n_obs=3000 with the NumPy/SciPy oracle behind "
                                              f"the class (threaded LAPACK; warm-started from the fixture's hyper-parameters)"}
        _emit(line)
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS))
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--pool", type=int, default=0, help="candidate pool size (default: the configuration's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep-mode", default="auto", choices=["auto", "fp64", "i8x7", "i8x8"],
                    help="variance contraction of the sweep (bo_set_sweep_mode); auto resolves on the global pool size")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
        return
    ctx = Ctx(args)
    {"C1": bench_c1, "C2": bench_sweep, "C3": bench_sweep, "C4": bench_c4, "C5": bench_c5}[args.config](ctx, args.config, cfg)
    ctx.finish()


if __name__ == "__main__":
    main()
