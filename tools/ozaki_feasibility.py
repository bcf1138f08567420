"""CPU numerics study for DESIGN §7: can the sweep's variance contraction u = L^-1 k* run as an error-free
(Ozaki scheme I) product of signed 7-bit slices on the INT8 tensor path and still meet the 1e-8-relative
variance bar?

Emulation (exact, no GPU): every row of L^-1 and every candidate column of K(X,X*) is scaled by a power of two
to |x| <= 1 and cut into S signed digits d_s in [-64, 64] with x ~= sum_s d_s 2^(-6-7s) (round-to-nearest
residual recursion, all steps exact in FP64).  Slice products d_s^A . d_t^B are integer dot products
(|sum| <= 64*64*n*S < 2^31 for n <= 8192, i.e. what an INT32 TMEM accumulator holds); slice pairs with the
same s+t share one accumulator; the groups are recombined in FP64 from the least significant one up.  Here the
integer products are float64 BLAS calls on integer-valued matrices (exact below 2^53).

Compared against: a longdouble (64-bit mantissa) evaluation of the same contraction (the truth) and the plain
FP64 BLAS product (what DMMA delivers today).  Reported: relative error of sigma^2 = k** - ||u||^2 over a
candidate set that includes points 1e-3..1e-5 away from training rows (sigma^2 << k**).

    python tools/ozaki_feasibility.py [n] [d] [cands]
"""
import os
import sys
import time

import numpy as np
import scipy.linalg as sla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gp_oracle as o  # noqa: E402  (tools/ may use the oracle as a checker, never the product)


def slices(x, S, W=7):
    """x (|x| <= 1) -> S integer-valued float64 arrays, x ~= sum d_s 2^(-(W-1)-W s).
    W=7: balanced digits in [-64, 64] (round to nearest), all slices signed int8.
    W=8: floor digits - the leading one signed in [-128, 127] (x pre-scaled to |x| <= 1/2), the others
    unsigned in [0, 255] (tcgen05 kind::i8 takes the signedness per operand per instruction)."""
    out = []
    r = x.copy()
    for s in range(S):
        sc = 2.0 ** ((W - 1) + W * s)
        dgt = np.rint(r * sc) if W == 7 else np.floor(r * sc)
        out.append(dgt)
        r = r - dgt / sc
    return out, r


def field_digits(x, S, W):
    """The kernels' digit geometry (sweep_i8.cuh, I8Dig): v = rint(x 2^F), F = 6 + W (S - 1); a 7-bit top digit in [-64, 64] and
    S - 1 balanced W-bit digits in [-2^(W-1), 2^(W-1) - 1], read as bit fields of v + bias.  Returns the S digit arrays
    (top first, integer-valued float64) -- an exact decomposition: x' = d0 2^-6 + sum_{s>=1} d_s 2^(-6 - W s) = v 2^-F."""
    F = 6 + W * (S - 1)
    v = np.rint(np.asarray(x, dtype=np.float64) * 2.0 ** F).astype(object)            # exact Python integers
    bias = sum((1 << (W - 1)) << (W * k) for k in range(S - 1))
    u = v + bias
    low = [np.array([(int(t) >> (W * k)) & ((1 << W) - 1) for t in u.ravel()], dtype=np.float64).reshape(np.shape(x)) - (1 << (W - 1))
           for k in range(S - 1)]
    top = np.array([int(t) >> (W * (S - 1)) for t in u.ravel()], dtype=np.float64).reshape(np.shape(x))
    return [top] + low[::-1]


def sliced_matmul_fields(A, B, S, W):
    """A @ B through the kernels' digit geometry: slice pairs s + t < S, one exact integer accumulator per s + t."""
    ea = pow2_scale(A, 1); eb = pow2_scale(B, 0)
    As = field_digits(A / ea, S, W); Bs = field_digits(B / eb, S, W)
    acc, imax = None, 0.0
    for g in range(S - 1, -1, -1):
        G = None
        for s in range(S):
            t = g - s
            if 0 <= t < S:
                P = As[s] @ Bs[t]
                G = P if G is None else G + P
        imax = max(imax, float(np.abs(G).max()))
        term = G * 2.0 ** (-12 - W * g)
        acc = term if acc is None else acc + term
    return acc * ea * eb, S * (S + 1) // 2, imax


def pow2_scale(v, axis):
    m = np.max(np.abs(v), axis=axis, keepdims=True)
    m = np.where(m > 0, m, 1.0)
    return np.exp2(np.ceil(np.log2(m)))


def sliced_matmul(A, B, S, triangular=True, W=7):
    ea = pow2_scale(A, 1) * (2.0 if W == 8 else 1.0)          # per row of A
    eb = pow2_scale(B, 0) * (2.0 if W == 8 else 1.0)          # per column of B
    As, ra = slices(A / ea, S, W)
    Bs, rb = slices(B / eb, S, W)
    ngroups = S if triangular else 2 * S - 1
    acc = None
    imax = 0.0
    for g in range(ngroups - 1, -1, -1):     # least significant group first
        G = None
        for s in range(S):
            t = g - s
            if 0 <= t < S:
                P = As[s] @ Bs[t]
                G = P if G is None else G + P
        imax = max(imax, float(np.abs(G).max()))
        term = G * 2.0 ** (-2 * (W - 1) - W * g)
        acc = term if acc is None else acc + term
    nprod = sum(1 for s in range(S) for t in range(S) if (s + t < S or not triangular))
    return acc * ea * eb, nprod, imax


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    C = int(sys.argv[3]) if len(sys.argv) > 3 else 128
    rng = np.random.default_rng(4)
    X = rng.random((n, d))
    y = np.sin(3.0 * X).sum(axis=1) + 0.05 * np.random.default_rng(5).standard_normal(n)
    y = (y - y.mean()) / y.std(ddof=1)
    t0 = time.time()
    gp = o.fit(X, y, o.KERNEL_MATERN52, 0.7, 1.0, 1e-3)
    Li = sla.solve_triangular(gp.L, np.eye(n), lower=True, check_finite=False)
    print(f"n={n} d={d}: fit + explicit inverse {time.time() - t0:.1f} s; max|L^-1|={np.abs(Li).max():.3g}, "
          f"median row max={np.median(np.abs(Li).max(axis=1)):.3g}, median |entry|={np.median(np.abs(Li[np.tril_indices(n)])):.3g}")
    far = rng.random((C // 2, d))
    near = []
    for j, eps in enumerate(np.logspace(-1.5, -5, C - C // 2)):
        near.append(np.clip(X[(j * 37) % n] + eps * rng.standard_normal(d), 0, 1))
    Xs = np.vstack([far, np.array(near)])
    Ks = o.kernel_matrix(gp.X, Xs, gp.kind, gp.lengthscale, gp.outputscale)          # (n, C)
    kss = o.prior_variance(Xs, gp.kind, gp.outputscale)

    t0 = time.time()
    Ul = np.tril(Li).astype(np.longdouble) @ Ks.astype(np.longdouble)
    var_true = kss.astype(np.longdouble) - np.einsum("ij,ij->j", Ul, Ul)
    print(f"longdouble truth {time.time() - t0:.1f} s; sigma^2 range {float(var_true.min()):.3e} .. {float(var_true.max()):.3e}")

    def report(name, U, extra=""):
        var = kss - np.einsum("ij,ij->j", U, U)
        rel = np.abs((var - var_true) / var_true).astype(np.float64)
        uerr = float(np.abs(U - Ul).max())
        print(f"  {name:<34s} max rel err sigma^2 = {rel.max():.2e}  median = {np.median(rel):.2e}  max|du| = {uerr:.2e} {extra}")
        return rel.max()

    report("FP64 BLAS (today's DMMA path)", Li @ Ks)
    for S in (6, 7, 8, 9):
        for tri in (True, False):
            if not tri and S not in (8,):
                continue
            t0 = time.time()
            U, nprod, imax = sliced_matmul(np.tril(Li), Ks, S, tri)
            report(f"S={S} slices, {'s+t<S' if tri else 'all pairs'} ({nprod} products)", U,
                   f"| max |int32 acc| = 2^{np.log2(imax):.1f}  [{time.time() - t0:.1f} s]")
    for S in (5, 6, 7):
        U, nprod, imax = sliced_matmul(np.tril(Li), Ks, S, True, W=8)
        report(f"radix 256: S={S}, s+t<S ({nprod} products)", U, f"| max |int32 acc| = 2^{np.log2(imax):.1f}")


if __name__ == "__main__":
    main()
