"""Objective test double with the reference simulator's duck type ("downward" seam).

``CachedCSVSimulator`` stands in for ``simulation.taichi.MPMSimulator`` (which needs Taichi and a display):
``configure_geometry(width, height)`` with the range validation of simulation/taichi.py:33-38,
``run_simulation(n, eta, sigma_y) -> float32[8]`` (simulation/taichi.py:46-62,140-142) answered from the
nearest cached results row, and ``cleanup()`` (simulation/taichi.py:145-148).  This is BASELINE.json's
config 1: "Taichi sim replaced by the cached CSV objective".
"""
from __future__ import annotations

import numpy as np

DEFAULT_BOUNDS = [(0.3, 1.0), (0.001, 300.0), (0.001, 400.0), (2.0, 7.0), (2.0, 7.0)]   # config/config.py:2-20


class CachedCSVSimulator:
    def __init__(self, params, outputs, bounds_list=DEFAULT_BOUNDS):
        self.params = np.asarray(params, dtype=np.float64)          # (N, 5) physical units
        self.outputs = np.asarray(outputs, dtype=np.float32)        # (N, 8)
        self.bounds = np.asarray(bounds_list, dtype=np.float64)
        span = self.bounds[:, 1] - self.bounds[:, 0]
        self._unit = (self.params - self.bounds[:, 0]) / span
        self._span = span
        self._geom = None
        self.calls = 0
        self.cleaned = False

    @classmethod
    def from_csv(cls, path, bounds_list=DEFAULT_BOUNDS):
        raw = np.genfromtxt(path, delimiter=",", skip_header=1, invalid_raise=False)
        raw = raw[~np.isnan(raw).any(axis=1)]
        return cls(raw[:, :5], raw[:, 5:13], bounds_list)

    def configure_geometry(self, width, height):
        (wlo, whi), (hlo, hhi) = self.bounds[3], self.bounds[4]
        if not (wlo <= width <= whi):
            raise ValueError(f"Width must be between {wlo} and {whi}")
        if not (hlo <= height <= hhi):
            raise ValueError(f"Height must be between {hlo} and {hhi}")
        self._geom = (float(width), float(height))

    def run_simulation(self, n, eta, sigma_y):
        for v, (lo, hi), name in zip((n, eta, sigma_y), self.bounds[:3], ("n", "eta", "sigma_y")):
            if not (lo <= v <= hi):
                raise ValueError(f"{name} must be between {lo} and {hi}")
        if self._geom is None:
            raise RuntimeError("configure_geometry must be called first")
        x = np.array([n, eta, sigma_y, self._geom[0], self._geom[1]], dtype=np.float64)
        u = (x - self.bounds[:, 0]) / self._span
        j = int(np.argmin(((self._unit - u) ** 2).sum(axis=1)))
        self.calls += 1
        return self.outputs[j].copy()

    def cleanup(self):
        self.cleaned = True
