"""-m gpu: randomised shapes (hypothesis) for the round-2 evaluation routes -- the error-bounded pruned kernel and the
symmetric LOO kernel against the reference-order brute-force kernels, the batched marginals against per-dimension
evaluation."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import kde_b200 as K
from tests.util import relerr

pytestmark = pytest.mark.gpu
SET = dict(max_examples=20, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)


def cloud(rng, d, N, clusters, spread):
    c = rng.normal(scale=3.0, size=(clusters, d))
    return (c[rng.integers(0, clusters, N)] + spread * rng.normal(size=(N, d))).T


@settings(**SET)
@given(d=st.integers(1, 6), N=st.integers(4100, 12000), M=st.integers(1, 5000), clusters=st.integers(1, 6),
       bw=st.floats(0.01, 2.0), seed=st.integers(0, 10 ** 6))
def test_bounded_eval_random_shapes(d, N, M, clusters, bw, seed):
    rng = np.random.default_rng(seed)
    pts = cloud(rng, d, N, clusters, 0.5)
    pos = np.hstack([cloud(rng, d, M, clusters, 0.7), rng.normal(scale=30.0, size=(d, 3))])   # a few far queries too
    p = K.kde(pts, np.full(d, bw), rng.random(N) + 0.01)
    brute = K.evaluateDualTree(p, pos)
    got = K.evaluateDualTree(p, pos, precision=K.F64_BOUNDED)
    assert np.array_equal(got == 0.0, brute == 0.0)
    nz = brute > 0
    assert relerr(got[nz], brute[nz]) < 1e-13
    assert relerr(K.evaluateDualTree(p, p, precision=K.F64_BOUNDED), K.evaluateDualTree(p, p)) < 1e-13


@settings(**SET)
@given(d=st.integers(1, 4), N=st.integers(4096, 20000), clusters=st.integers(1, 5), bw=st.floats(0.005, 3.0),
       seed=st.integers(0, 10 ** 6))
def test_symmetric_loo_random_shapes(d, N, clusters, bw, seed):
    rng = np.random.default_rng(seed)
    pts = cloud(rng, d, N, clusters, 0.5)
    p = K.kde(pts, np.full(d, bw), rng.random(N) + 0.01)
    try:
        K.set_pruning(0)
        H0 = K.entropy(p)
        K.set_pruning(1)
        H1 = K.entropy(p)
    finally:
        K.set_pruning(1)
    assert (np.isinf(H0) and np.isinf(H1)) or abs(H1 - H0) <= 2e-13 * abs(H0)


@settings(**SET)
@given(d=st.integers(1, 8), N=st.integers(1, 3000), G=st.integers(1, 300), bw=st.floats(0.05, 2.0), seed=st.integers(0, 10 ** 6))
def test_marginals_random_shapes(d, N, G, bw, seed):
    rng = np.random.default_rng(seed)
    pts = rng.normal(size=(d, N))
    p = K.kde(pts, np.full(d, bw), rng.random(N) + 0.01)
    X = rng.normal(scale=1.5, size=(d, G))
    Y = K.eval_marginals(p, X)
    k = int(rng.integers(0, d))
    ref = K.evaluateDualTree(K.kde(pts[k:k + 1], [bw], K.getWeights(p)), X[k:k + 1])
    assert relerr(Y[k], ref) < 1e-12
