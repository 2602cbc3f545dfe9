import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib_built():
    from mtamrecommender_b200 import build
    return build.build()


# ---- the conditioning allowance of the parity tests -------------------------------------------------------------------
# Gradients and Adam-updated weights are held to 1e-4 (norm-wise, per tensor) against the fp64 oracle.  Where the graph
# itself is ill-conditioned in fp32 on the test input (a ReLU pre-activation within rounding of its kink), the oracle's
# own fp32 evaluation misses that bar too; such a tensor may use 3x the oracle's fp32-vs-fp64 error instead, CAPPED at
# 1e-3, and every use is listed in the terminal summary.
TOL_BASE, TOL_CAP = 1e-4, 1e-3
_allowance_used = []


def parity_tol(test: str, tensor: str, oracle_fp32_err: float) -> float:
    tol = min(max(TOL_BASE, 3.0 * oracle_fp32_err), TOL_CAP)
    if tol > TOL_BASE:
        _allowance_used.append((test, tensor, oracle_fp32_err, tol))
    return tol


def pytest_terminal_summary(terminalreporter):
    if _allowance_used:
        terminalreporter.write_line(f"parity: {len(_allowance_used)} tensor comparisons used the conditioning allowance "
                                    f"(tolerance above {TOL_BASE:g}, capped at {TOL_CAP:g}):")
        for test, tensor, err, tol in _allowance_used:
            terminalreporter.write_line(f"  {test}: {tensor}: oracle fp32-vs-fp64 {err:.2e} -> tolerance {tol:.2e}")
