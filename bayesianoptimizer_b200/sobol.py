"""Host side of the in-kernel scrambled-Sobol candidate pool.

The state (direction numbers + digital shift) is taken from a fresh
``torch.quasirandom.SobolEngine(d, scramble=True, seed)`` so that candidate ``i`` generated on the
device equals row ``i`` of ``SobolEngine.draw(dtype=float64)`` bit for bit.  Stands in for the host-side
pool construction of optimization/Bayesian7.py:650-655 and the raw samples of ``optimize_acqf``
(optimization/Bayesian.py:105-112).
"""
from __future__ import annotations

import torch

from ._lib import BO_MAX_DIM, BO_SOBOL_BITS, BoSobol


def sobol_state(d: int, seed: int) -> BoSobol:
    if not 1 <= d <= BO_MAX_DIM:
        raise ValueError(f"d must be in [1, {BO_MAX_DIM}]")
    eng = torch.quasirandom.SobolEngine(d, scramble=True, seed=int(seed))
    if eng.MAXBIT != BO_SOBOL_BITS:
        raise RuntimeError("SobolEngine.MAXBIT changed; the in-kernel generator assumes 30 bits")
    state = eng.sobolstate.tolist()
    shift = eng.shift.tolist()
    st = BoSobol()
    st.d = d
    for k in range(d):
        for b in range(BO_SOBOL_BITS):
            st.direction[k][b] = int(state[k][b])
        st.shift[k] = int(shift[k])
    return st
