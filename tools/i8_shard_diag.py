"""Diagnostic: are per-candidate values of a pinned sliced mode independent of the shard layout / repeatable run to run?"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from bayesianoptimizer_b200 import GPEngine, sobol_state
from conftest import synth_problem
mode = sys.argv[1] if len(sys.argv) > 1 else "i8x7"
X, y = synth_problem(384, 6, 21, 22)
eng = GPEngine(torch.device("cuda", 0))
eng.fit(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(), "matern52", 0.7, 1.0, 1e-3)
eng.set_sweep_mode(mode)
st = sobol_state(6, 17)
N = 50_000
bf = float(y.max())
def run(lo, cnt):
    v, i, m, var, a = eng.sweep("ei", bf, sobol=st, first_index=lo, count=cnt, topk=8, return_all=True)
    return m.cpu().numpy(), var.cpu().numpy(), a.cpu().numpy(), eng.last_sweep_flagged()
m1, v1, a1, f1 = run(0, N)
m2, v2, a2, f2 = run(0, N)
print(mode, "repeat: mean equal", np.array_equal(m1, m2), "var equal", np.array_equal(v1, v2), "acq equal", np.array_equal(a1, a2), "flagged", f1, f2)
for G in (2, 3, 13):
    per = -(-N // G)
    ms, vs = [], []
    for r in range(G):
        lo = r * per; cnt = max(0, min(per, N - lo))
        m, v, a, f = run(lo, cnt)
        ms.append(m); vs.append(v)
    m = np.concatenate(ms); v = np.concatenate(vs)
    dm = np.nonzero(m != m1)[0]; dv = np.nonzero(v != v1)[0]
    print(f"G={G}: mean differs at {len(dm)} candidates, var differs at {len(dv)}; first {dv[:5].tolist()} var {v1[dv[:3]]} vs {v[dv[:3]]}; positions mod 64: {(dv[:8] % 64).tolist()} shard-local mod 64: {((dv[:8] % per) % 64).tolist()}")
eng.close()
