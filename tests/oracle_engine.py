"""Test double with GPEngine's interface, backed by the CPU oracle.  TESTS ONLY: lets the host-side logic
of BayesianOptimizer (CSV / resume / LHS / batch loop / sharding plumbing) run in the CPU tier, and serves
as the reference in GPU-tier comparisons of the optimizer loop.  The product never imports this."""
import numpy as np
import torch

from oracle import gp_oracle as o

_K = {"matern52": o.KERNEL_MATERN52, "rbf": o.KERNEL_RBF, "linear_matern52": o.KERNEL_LINEAR_MATERN52}
_A = {"ei": o.ACQ_EI, "logei": o.ACQ_LOGEI, "ucb": o.ACQ_UCB, "var": o.ACQ_VAR, "mean": o.ACQ_MEAN}


def _sobol_arrays(sob):
    d = sob.d
    st = np.array([[sob.direction[k][b] for b in range(30)] for k in range(d)], dtype=np.int64)
    sh = np.array([sob.shift[k] for k in range(d)], dtype=np.int64)
    return st, sh


class OracleEngine:
    device = torch.device("cpu")

    sweep_mode = "fp64"

    def __init__(self):
        self.gp = None
        self.svgp = None
        self.n = self.d = 0
        self.fits = 0

    # the oracle has one (FP64) contraction; the mode interface of GPEngine is accepted and ignored
    def set_sweep_mode(self, mode="auto"):
        self.sweep_mode = mode

    def resolve_sweep_mode(self, pool_total):
        return "fp64"

    def fit(self, X, y, kernel="matern52", lengthscale=1.0, outputscale=1.0, noise=1e-3, mean=0.0, jitter=0.0,
            linear_variance=0.0):
        from bayesianoptimizer_b200 import NotPositiveDefiniteError
        X = torch.as_tensor(X).cpu().numpy()
        y = torch.as_tensor(y).cpu().numpy().reshape(-1)
        try:
            self.gp = o.fit(X, y, _K[kernel], lengthscale, outputscale, noise, mean, jitter, linear_variance)
        except o.NotPositiveDefinite as e:
            raise NotPositiveDefiniteError(e.pivot)
        self.n, self.d = X.shape
        self.fits += 1
        return self

    def load_svgp(self, Z, var_mean, var_chol, kernel="linear_matern52", lengthscale=1.0, outputscale=1.0,
                  linear_variance=0.0, mean=0.0, noise=0.0, jitter=1e-6):
        npy = lambda a: torch.as_tensor(a, dtype=torch.float64).cpu().numpy()
        self.svgp = o.SVGPTask(npy(Z), _K[kernel], npy(lengthscale).reshape(-1), float(outputscale), float(linear_variance),
                               float(mean), float(noise), float(jitter), npy(var_mean).reshape(-1), npy(var_chol))
        self.n, self.d = self.svgp.Z.shape
        return self

    def posterior(self, Xs, min_variance=1e-6):
        if getattr(self, "svgp", None) is not None:
            mu, var = o.svgp_predict(self.svgp, torch.as_tensor(Xs).cpu().numpy(), min_variance)
        else:
            mu, var = o.posterior(self.gp, torch.as_tensor(Xs).cpu().numpy(), min_variance)
        return torch.from_numpy(mu), torch.from_numpy(var)

    def posterior_multi(self, Y, Xs, means=None, min_variance=1e-6, with_variance=True):
        mu, var = o.posterior_multi(self.gp, torch.as_tensor(Y).cpu().numpy(), torch.as_tensor(Xs).cpu().numpy(),
                                    None if means is None else np.asarray(means, dtype=np.float64), min_variance)
        return torch.from_numpy(mu), (torch.from_numpy(var) if with_variance else None)

    def sweep(self, acq="ei", best_f=0.0, beta=2.0, candidates=None, sobol=None, first_index=0, count=None, topk=1,
              min_variance=1e-6, return_all=False):
        if candidates is not None:
            pts = torch.as_tensor(candidates).cpu().numpy().reshape(-1, self.d)
        else:
            pts = o.sobol_points(*_sobol_arrays(sobol), first_index, count)
        if getattr(self, "svgp", None) is not None:                       # SVGP state: variance / mean scores only
            mu, var = o.svgp_predict(self.svgp, pts, min_variance)
            av = var if acq == "var" else mu
            tv, ti = o.topk(av, topk, first_index) if topk else (np.empty(0), np.empty(0, dtype=np.int64))
            out = (torch.from_numpy(np.asarray(tv, dtype=np.float64)), torch.from_numpy(np.asarray(ti, dtype=np.int64)))
            return out + ((torch.from_numpy(mu), torch.from_numpy(var), torch.from_numpy(av)) if return_all else ())
        tv, ti, mu, var, av = o.sweep(self.gp, pts, _A[acq], best_f, beta, k=topk, first_index=first_index,
                                      min_variance=min_variance)
        vals = np.full(topk, -np.inf); idx = np.full(topk, -1, dtype=np.int64)
        vals[:len(tv)] = tv; idx[:len(ti)] = ti
        out = (torch.from_numpy(vals), torch.from_numpy(idx))
        if return_all:
            out += (torch.from_numpy(mu), torch.from_numpy(var), torch.from_numpy(av))
        return out

    def sobol_points(self, sobol, idx):
        st, sh = _sobol_arrays(sobol)
        idx = torch.as_tensor(idx).cpu().numpy().reshape(-1)
        return torch.from_numpy(np.vstack([o.sobol_points(st, sh, int(i), 1) for i in idx]).reshape(-1, sobol.d))

    def refine(self, starts, acq="ei", best_f=0.0, beta=2.0, iters=50, min_variance=1e-6):
        import scipy.optimize as so
        starts = torch.as_tensor(starts).cpu().numpy().reshape(-1, self.d)
        xs, vs = [], []
        for s in starts:
            r = so.minimize(lambda z: tuple(-np.asarray(t) for t in o.acquisition_with_grad(self.gp, z, _A[acq], best_f, beta)),
                            s, jac=True, method="L-BFGS-B", bounds=[(0.0, 1.0)] * self.d, options={"maxiter": iters})
            xs.append(r.x); vs.append(-r.fun)
        return torch.from_numpy(np.array(xs)), torch.from_numpy(np.array(vs))

    def append(self, x, y=None):
        from bayesianoptimizer_b200 import NotPositiveDefiniteError
        try:
            self.gp = o.append_point(self.gp, torch.as_tensor(x).cpu().numpy(), y)
        except o.NotPositiveDefinite as e:
            raise NotPositiveDefiniteError(e.pivot)
        self.n += 1
        return self

    def lml_grad_batched(self, X, y, thetas, kernel="matern52", mean=0.0):
        X = torch.as_tensor(X).cpu().numpy(); y = torch.as_tensor(y).cpu().numpy().reshape(-1)
        d = X.shape[1]
        lin = _K[kernel] == o.KERNEL_LINEAR_MATERN52
        p = d + 3 if lin else d + 2
        th = np.asarray(thetas, dtype=np.float64).reshape(-1, p)
        lml, grad, status = [], [], []
        for t in th:
            try:
                l, g = o.lml_and_grad(X, y, _K[kernel], np.exp(t[:d]), np.exp(t[d]), np.exp(t[d + 1]), mean,
                                      np.exp(t[d + 2]) if lin else 0.0)
                lml.append(l); grad.append(g); status.append(0)
            except o.NotPositiveDefinite as e:
                lml.append(-np.inf); grad.append(np.zeros(p)); status.append(e.pivot)
        return torch.tensor(lml), torch.from_numpy(np.array(grad)), torch.tensor(status, dtype=torch.int32)

    def topk_scores(self, scores, k, first_index=0):
        tv, ti = o.topk(torch.as_tensor(scores).cpu().numpy().reshape(-1), int(k), first_index)
        vals = np.full(int(k), -np.inf); idx = np.full(int(k), -1, dtype=np.int64)
        vals[:len(tv)] = tv; idx[:len(ti)] = ti
        return torch.from_numpy(vals), torch.from_numpy(idx)

    def fps(self, X, m, start=0):
        return torch.from_numpy(o.fps(torch.as_tensor(X).cpu().numpy(), m, start))

    def close(self):
        pass
