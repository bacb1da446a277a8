"""Shared synthetic inputs for the parity tests (seeded, identical for oracle and CUDA path)."""
import numpy as np


def mixture(rng, d, N, shift=0.0, sigma=0.6):
    """K=4 isotropic Gaussians on the first 4 corners of {+-2}^d."""
    corners = np.array([[(-2.0 if (c >> (d - 1 - k)) & 1 == 0 else 2.0) for k in range(d)] for c in range(min(4, 2 ** d))])
    comp = rng.integers(0, len(corners), size=N)
    pts = corners[comp].T + sigma * rng.standard_normal((d, N))
    pts[0, :] += shift
    return pts


def silverman(pts):
    d, N = pts.shape
    return pts.std(axis=1, ddof=1) * (4.0 / ((d + 2.0) * N)) ** (1.0 / (d + 4.0))


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.abs(a - b) / den)) if a.size else 0.0
