"""Committed golden vectors (tests/golden/small_window.npz, generated from the oracle by tests/golden/make_golden.py):
the oracle must keep reproducing them (CPU), the CUDA engine must match them through the C-ABI (GPU)."""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(spec)
spec.loader.exec_module(make_golden)
GOLD = dict(np.load(os.path.join(HERE, "golden", "small_window.npz")))


def _compare(got, rtol_exact, solve_tol):
    for k, ref in GOLD.items():
        v = np.asarray(got[k])
        if ref.dtype.kind in "iu" or k in ("solve_iterations", "prior_n"):
            assert np.array_equal(v, ref), k
        elif k.startswith("solve_") and k not in ("solve_initial_cost",):
            scale = max(1.0, np.abs(ref).max())
            assert np.allclose(v, ref, rtol=0, atol=solve_tol * scale), (k, np.abs(v - ref).max())
        elif k.startswith("prior_"):
            scale = np.abs(ref).max()
            assert np.allclose(v, ref, rtol=0, atol=1e-7 * scale), (k, np.abs(v - ref).max() / scale)
        else:
            scale = max(1e-300, np.abs(ref).max())
            assert np.allclose(v, ref, rtol=0, atol=rtol_exact * scale), (k, np.abs(v - ref).max() / scale)


def test_oracle_reproduces_golden(oracle_lib):
    _compare(make_golden.golden_cases(oracle_lib), rtol_exact=1e-13, solve_tol=1e-10)


@pytest.mark.gpu
def test_cuda_matches_golden(cuda_lib):
    # factor values / Jacobians / normal equations: different summation orders only; solve: north-star tolerance
    _compare(make_golden.golden_cases(cuda_lib), rtol_exact=1e-9, solve_tol=1e-5)
