"""-m gpu: K6, the error-bounded tile-pruned evaluation (csrc/eval_pruned.cu) -- the B200 counterpart of the reference's
dual-tree evaluation (src/DualTree01.jl:164-299, switched on by setForceEvalDirect!(false)).

Contract under test: every value is within 1e-13 relative of the brute-force sum (so within the 1e-12 parity bar against
the oracle), exact zeros / subnormals keep the reference's behaviour, and something is actually pruned on clustered
data.  The reference's own dual tree is only within errTol = 1e-3 of brute force."""
import math

import numpy as np
import pytest

import kde_b200 as K
from oracle.oracle import OKDE
from tests.util import mixture, relerr, silverman

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def default_policy():
    K.set_pruning(1)
    yield
    K.set_pruning(1)


@pytest.mark.parametrize("d,N,M", [(1, 20_000, 30_001), (2, 9_999, 5_000), (3, 50_000, 20_003), (5, 6_000, 4_100), (8, 5_000, 3_000)])
def test_bounded_eval_matches_brute_force_and_oracle(d, N, M):
    rng = np.random.default_rng(10 * d + N)
    pts, pos = mixture(rng, d, N), mixture(rng, d, M) * 1.1
    w = rng.random(N) + 0.05
    bw = silverman(pts) * 0.5
    p = K.kde(pts, bw, w)
    brute = K.evaluateDualTree(p, pos)
    got = K.evaluateDualTree(p, pos, precision=K.F64_BOUNDED)
    kept, redo = K.pruned_stats()
    assert relerr(got, brute) < 1e-13
    assert kept < (0.9 if d <= 3 else 1.0001), kept       # clustered data: far cluster pairs are dropped
    o = OKDE.kde_bw(pts, bw, w)
    assert relerr(got[:200], o.evaluate(pos[:, :200], nthreads=8)) < 1e-12
    loo_b, loo_p = K.evaluateDualTree(p, p), K.evaluateDualTree(p, p, precision=K.F64_BOUNDED)
    assert relerr(loo_p, loo_b) < 1e-13
    perm = p.bt.permutation[N:] - 1
    assert relerr(loo_p[perm[100:300]], o.loo_rows(100, 300, nthreads=8)) < 1e-12


def test_far_queries_take_the_exact_pass_and_keep_the_zero_rule():
    rng = np.random.default_rng(3)
    pts = mixture(rng, 2, 8000)
    p = K.kde(pts, [0.05, 0.05])
    far = np.array([[50.0, -3.0, 0.0, 2.05], [50.0, 40.0, 0.0, -2.0]])          # two hopeless points, two ordinary ones
    pos = np.hstack([mixture(rng, 2, 3000), far])
    brute = K.evaluateDualTree(p, pos)
    got = K.evaluateDualTree(p, pos, precision=K.F64_BOUNDED)
    kept, redo = K.pruned_stats()
    assert redo >= 2                                        # the far points went through the all-components pass
    assert np.array_equal(got == 0.0, brute == 0.0) and got[-4] == 0.0
    nz = brute > 0
    assert relerr(got[nz], brute[nz]) < 1e-13
    o = OKDE.kde_bw(pts, [0.05, 0.05])
    exp = o.evaluate(pos[:, -4:])
    assert np.array_equal(exp == 0.0, got[-4:] == 0.0) and relerr(got[-2:], exp[-2:]) < 1e-12


def test_loo_likelihood_uses_the_bounded_kernel_and_agrees_with_brute_force():
    rng = np.random.default_rng(4)
    N = 40_000
    x = mixture(rng, 1, N)
    p = K.kde(x, [0.01])
    K.set_pruning(0)
    H0 = K.entropy(p)
    K.set_pruning(1)
    H1 = K.entropy(p)
    kept, _ = K.pruned_stats()
    assert kept < 0.5 and abs(H1 - H0) <= 1e-13 * abs(H0)
    o = OKDE.kde_bw(x, [0.01])
    rows = o.loo_rows(0, 256, nthreads=8)
    L = K.evaluateDualTree(p, p, precision=K.F64_BOUNDED)
    assert relerr(L[p.bt.permutation[N:N + 256] - 1], rows) < 1e-12
    # an isolated point far from everything: its LOO density underflows to exactly 0 => +Inf entropy, both routes
    q = K.kde(np.hstack([x, [[1e4]]]), [0.01])
    assert K.entropy(q) == math.inf
    K.set_pruning(0)
    assert K.entropy(q) == math.inf
    # wide bandwidth: nothing can be dropped, the brute-force kernel serves the call (same bits as mode 0)
    K.set_pruning(1)
    pw = K.kde(x, [3.0])
    a = K.entropy(pw)
    K.set_pruning(0)
    assert a == K.entropy(pw)


def test_kde_lcv_with_and_without_pruning_selects_the_same_bandwidths():
    rng = np.random.default_rng(5)
    pts = mixture(rng, 2, 30_000)
    K.set_pruning(0)
    b0 = K.lcv_bandwidths(pts)
    K.set_pruning(1)
    b1 = K.lcv_bandwidths(pts)
    assert relerr(b1, b0) < 1e-9


def test_set_force_eval_direct_mirror():
    """setForceEvalDirect!(false) (src/DualTree01.jl:3-9): evaluateDualTree switches to the pruned route."""
    rng = np.random.default_rng(6)
    pts, pos = mixture(rng, 3, 20_000), mixture(rng, 3, 5000)
    p = K.kde(pts, silverman(pts) * 0.4)
    brute = K.evaluateDualTree(p, pos)
    K.setForceEvalDirect(False)
    got = K.evaluateDualTree(p, pos)
    kept, _ = K.pruned_stats()
    K.setForceEvalDirect(True)
    assert kept < 0.9 and relerr(got, brute) < 1e-13 and np.array_equal(K.evaluateDualTree(p, pos), brute)
    with pytest.raises(K.KDEError):
        K.set_pruning(7)


def test_bounded_eval_sharded_in_process():
    K.init_multi(devices=[0, 0, 0])
    try:
        rng = np.random.default_rng(7)
        pts, pos = mixture(rng, 3, 30_000), mixture(rng, 3, 40_001)
        p = K.kde(pts, silverman(pts) * 0.5)
        got = K.evaluateDualTree(p, pos, precision=K.F64_BOUNDED)
        loo = K.evaluateDualTree(p, p, precision=K.F64_BOUNDED)
        K.init_multi(1)
        assert relerr(got, K.evaluateDualTree(p, pos)) < 1e-13 and relerr(loo, K.evaluateDualTree(p, p)) < 1e-13
    finally:
        K.init_multi(1)


@pytest.mark.parametrize("d,N,scale", [(1, 5000, 0.3), (1, 100_000, 1.0), (2, 12_345, 0.5), (3, 9_000, 1.0), (4, 7_001, 1.0), (7, 4_500, 1.0)])
def test_symmetric_loo_likelihood_matches_reference_order_kernel_and_oracle(d, N, scale):
    """The each-pair-once LOO kernel (loo_sym_kernel) that serves entropy / nLOO_LL for 4096 <= N <= 262144 against
    the reference-order brute-force kernel (kdeb200_set_pruning(0)) and the oracle's literal rows: non-uniform
    weights, ragged sizes (row blocks, tiles and N all out of step), bandwidths with and without anything to prune."""
    rng = np.random.default_rng(100 * d + N)
    pts = mixture(rng, d, N)
    w = rng.random(N) + 0.05
    bw = silverman(pts) * scale
    p = K.kde(pts, bw, w)
    K.set_pruning(0)
    H0 = K.entropy(p)
    K.set_pruning(1)
    H1 = K.entropy(p)
    assert abs(H1 - H0) <= 2e-13 * abs(H0), (H0, H1)
    for a in (0.3, 2.5):                                   # nLOO_LL multiplies the bandwidth: narrow and wide kernels
        K.set_pruning(0)
        r0 = K.nLOO_LL(a, p)
        K.set_pruning(1)
        r1 = K.nLOO_LL(a, p)
        assert abs(r1 - r0) <= 2e-13 * abs(r0)
    if N <= 20_000:
        o = OKDE.kde_bw(pts, bw, w)
        assert abs(H1 - o.entropy()) <= 1e-12 * abs(H1)


@pytest.mark.parametrize("d,N,M", [(1, 30_000, 100_001), (3, 50_000, 80_000), (2, 9_999, 777), (6, 6_000, 5_000)])
def test_fp32_bounded_eval_within_1e5(d, N, M):
    """KDEB200_F32_BOUNDED: the packed-FP32 kernel through the box-pair pruning (window 8.6 bandwidths, rows below the
    bound recomputed in FP64) keeps the FP32 contract of 1e-5 against the FP64 brute-force sum; far queries keep their
    exact zeros / tiny values through the FP64 exact pass."""
    rng = np.random.default_rng(7 * d + N)
    pts, pos = mixture(rng, d, N), mixture(rng, d, M) * 1.05
    p = K.kde(pts, silverman(pts) * 0.5, rng.random(N) + 0.05)
    pos[:, :3] = 40.0                                      # hopeless queries: exact pass
    ref = K.evaluateDualTree(p, pos)
    got = K.evaluateDualTree(p, pos, precision=K.F32_BOUNDED)
    kept, redo = K.pruned_stats()
    assert redo >= 3 and kept < (0.95 if d <= 3 else 1.0001)
    nz = ref > 1e-300
    assert relerr(got[nz], ref[nz]) < 1e-5 and np.array_equal(got[~nz] == 0.0, ref[~nz] == 0.0)
    # leave-one-out calls run the unpruned FP32 kernel: its 1e-5 holds where the sum is not dominated by far-tail terms
    # (FP32 exponents of ~100 carry ~1e-5 of relative error by themselves)
    l32, l64 = K.evaluateDualTree(p, p, precision=K.F32_BOUNDED), K.evaluateDualTree(p, p)
    body = l64 > 1e-3 * np.median(l64)
    assert body.mean() > 0.9 and relerr(l32[body], l64[body]) < 1e-5
