"""Which position attribute of the CTA-pair kernel changes a candidate's bits? Sweep the same candidates at shifted first_index."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from bayesianoptimizer_b200 import GPEngine, sobol_state
from conftest import synth_problem
X, y = synth_problem(384, 6, 21, 22)
eng = GPEngine(torch.device("cuda", 0))
eng.fit(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(), "matern52", 0.7, 1.0, 1e-3)
eng.set_sweep_mode("i8x8")
st = sobol_state(6, 17)
bf = float(y.max())
def run(lo, cnt):
    v, i, m, var, a = eng.sweep("ei", bf, sobol=st, first_index=lo, count=cnt, topk=8, return_all=True)
    return m.cpu().numpy(), var.cpu().numpy()
m0, v0 = run(0, 20000)
for sh in (1, 2, 4, 8, 16, 32, 64, 128, 9472):
    m, v = run(sh, 20000 - sh)
    dm = (m != m0[sh:]).sum(); dv = (v != v0[sh:]).sum()
    print(f"shift {sh:5d}: mean differs {dm:6d}, var differs {dv:6d} of {20000 - sh}")
eng.close()
eng = GPEngine(torch.device("cuda", 0))
eng.fit(torch.from_numpy(X).cuda(), torch.from_numpy(y).cuda(), "matern52", 0.7, 1.0, 1e-3)
eng.set_sweep_mode("i8x8")
m0, v0 = run(0, 20000)
for sh in (8, 64):
    m, v = run(sh, 20000 - sh)
    print("shift", sh, "differing local indices (mean):", np.nonzero(m != m0[sh:])[0].tolist()[:80])
    print("   var:", np.nonzero(v != v0[sh:])[0].tolist()[:80])
m, v = run(0, 19990)
print("same start, N=19990: mean differs at", np.nonzero(m != m0[:19990])[0].tolist()[:40])
m, v = run(0, 20000)
print("repeat: ", (m != m0).sum(), (v != v0).sum())
eng.close()
