"""Shared helpers for the parity tests."""
import numpy as np

REL_TOL = 1e-9  # BASELINE.json north_star: FP64 state and covariance within 1e-9 relative


def rel_err_state(x, ref):
    """max over filters of |x - ref| / max(1, |ref|) (positions are O(1..10) m)."""
    return float(np.max(np.abs(x - ref) / np.maximum(1.0, np.abs(ref))))


def rel_err_cov(P, ref):
    """per-filter matrix-norm relative error: max|P - ref| / max|ref| over each filter's entries."""
    num = np.abs(P - ref).max(axis=0)
    den = np.maximum(np.abs(ref).max(axis=0), 1e-300)
    return float(np.max(num / den))
