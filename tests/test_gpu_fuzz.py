"""GPU tier: randomised parity sweep (tools/fuzz_parity.py) -- random n / d / kernel kind / hyper-parameters / pools /
acquisitions / appends / SVGP states / batched LML / large top-K against the CPU oracle.  1800 cases were run clean during
development (seeds 1-3 x 600); the suite keeps a short slice."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_randomised_parity_slice():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_parity.py"), "80", "11"], capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "80 cases, 0 failures" in r.stdout


def test_randomised_parity_slice_on_the_sliced_sweep():
    """The same generator with the contraction pinned to the INT8-sliced path (8 slices): every eligible model (exact GP,
    Matern / RBF, n > 128) takes it whatever the pool size; 160 cases ran clean during development (seed 21,
    profiles/r01_fuzz_i8x8.log) -- the suite keeps their first 40."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_parity.py"), "40", "21", "i8x8"], capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "40 cases, 0 failures" in r.stdout
