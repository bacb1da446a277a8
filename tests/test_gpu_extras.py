"""-m gpu: the batched callers around the evaluation seam (csrc/extras.cu): every marginal on its
grid in one launch (getKDEMax), the whole N x N grid of intersIntgAppxIS in one call per density, and sample / rand /
resample with the component draw and the kernel perturbation on the device."""
import numpy as np
import pytest

import kde_b200 as K
from oracle.oracle import OKDE
from tests.util import mixture, relerr, silverman

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("d,N,G", [(1, 300, 200), (3, 5000, 200), (4, 777, 33), (8, 100, 5)])
def test_eval_marginals_equals_marginal_trees(d, N, G):
    """kdeb200_eval_marginals vs the reference's route marginal(p,[i]) -> evaluate (oracle), 1e-12."""
    rng = np.random.default_rng(d * 100 + N)
    pts = mixture(rng, d, N)
    w = rng.random(N) + 0.1
    bw = silverman(pts)
    p, o = K.kde(pts, bw, w), OKDE.kde_bw(pts, bw, w)
    X = np.stack([np.linspace(pts[k].min() - 1, pts[k].max() + 1, G) for k in range(d)])
    Y = K.eval_marginals(p, X)
    for k in range(d):
        om = o.marginal([k + 1])
        assert relerr(Y[k], om.evaluate(X[k].reshape(1, G))) < 1e-12
        assert relerr(Y[k], K.evaluateDualTree(K.marginal(p, [k + 1]), X[k].reshape(1, G))) < 1e-12
    far = X.copy()
    far[:, 0] = 1e6                                        # exact zeros in the far tail, like the evaluation kernel
    assert np.all(K.eval_marginals(p, far)[:, 0] == 0.0)
    with pytest.raises(K.KDEError):
        K.eval_marginals(p, np.zeros((d + 1, 4)))


def test_get_kde_max_matches_the_per_dimension_route():
    rng = np.random.default_rng(5)
    p = K.kde(3.0 + 0.5 * rng.standard_normal((3, 2000)))
    m = K.getKDEMax(p, N=200)
    ref = np.zeros(3)
    for i in range(3):                                     # src/DualTree01.jl:558-569 literally
        mm = K.marginal(p, [i + 1])
        r = K.getKDERange(mm).ravel(order="F")
        X = np.linspace(r[0], r[1], 200)
        ref[i] = X[int(np.argmax(mm(X.reshape(1, 200))))]
    assert np.allclose(m, ref, rtol=0, atol=1e-12) and np.all(np.abs(m - 3.0) < 0.3)


def test_inters_intg_2d_grid_in_one_call_equals_row_by_row():
    rng = np.random.default_rng(6)
    p, q = K.kde(rng.standard_normal((2, 100))), K.kde(rng.standard_normal((2, 150)) + 0.3)
    N = 41
    got = K.intersIntgAppxIS(p, q, N=N)
    LD = [K.getKDERangeLinspace(K.marginal(p, [k + 1]), N=N, extend=0.3) for k in range(2)]
    dx = [ld[1] - ld[0] for ld in LD]
    xx = np.zeros((2, N))
    xx[0] = LD[0]
    acc = 0.0
    for i in range(N):                                     # src/DualTree01.jl:605-613
        xx[1, :] = LD[1][i]
        acc += (dx[0] * float(np.sum(K.evaluateDualTree(p, xx) * K.evaluateDualTree(q, xx)))) * dx[1]
    assert abs(got - acc) <= 1e-13 * abs(acc) and 0.03 < got < 0.2


def test_sample_with_injected_variates_is_the_reference_scan():
    """src/KDE01.jl:164-183 restated in numpy on the same uniforms / normals: identical component indices, points equal."""
    rng = np.random.default_rng(7)
    N, Np, d = 500, 4000, 3
    pts = mixture(rng, d, N)
    w = rng.random(N) + 0.01
    w[10:20] = 0.0                                         # zero-weight components are never drawn
    p = K.kde(pts, [0.2, 0.3, 0.4], w)
    U, G = rng.random(Np), rng.standard_normal((d, Np))

    class Inj:                                             # feeds the mirror's rng calls with the fixed arrays
        def standard_normal(self, shape):
            return G
        def random(self, n):
            return U
    got, idx = K.sample(p, Np, rng=Inj())
    cw = np.cumsum(K.getWeights(p))
    cw = cw / cw[-1]
    t = np.concatenate([np.sort(U), [10.0]])
    ref_pts, ref_idx, ii = np.zeros((d, Np)), np.zeros(Np, dtype=np.int64), 0
    P, B = K.getPoints(p), K.getBW(p)
    for i in range(N):
        while cw[i] > t[ii]:
            ref_pts[:, ii] = P[:, i] + B[:, i] * G[:, ii]
            ref_idx[ii] = i + 1
            ii += 1
    assert ii == Np and np.array_equal(idx, ref_idx) and np.max(np.abs(got - ref_pts)) < 1e-14
    assert not np.any((idx >= 11) & (idx <= 20))


def test_sample_rand_resample_statistics_and_seeding():
    rng = np.random.default_rng(8)
    pts = np.hstack([rng.normal(-3, 0.3, (1, 300)), rng.normal(2, 0.5, (1, 700))])
    p = K.kde(pts, [0.1])
    a, ia = K.sample(p, 50_000, seed=1)
    b, ib = K.sample(p, 50_000, seed=1)
    c, _ = K.sample(p, 50_000, seed=2)
    assert np.array_equal(a, b) and np.array_equal(ia, ib) and not np.array_equal(a, c)
    assert np.all(np.diff(ia) >= 0)                        # sorted uniforms => non-decreasing component indices
    assert abs(np.mean(a < -1) - 0.3) < 0.01 and abs(a.mean() - pts.mean()) < 0.03
    assert abs(a.std() - np.sqrt(pts.var() + 0.01)) < 0.03
    assert K.rand(p, 7, seed=3).shape == (1, 7)
    r = K.resample(p, 2000, seed=4)
    assert K.Npts(r) == 2000 and abs(K.getPoints(r).mean() - pts.mean()) < 0.15
    e, ie = K.sample(p, 5, ind=[1, 2, 3, 4, 5], rng=np.random.default_rng(0))
    assert e.shape == (1, 5) and np.array_equal(ie, [1, 2, 3, 4, 5])
    with pytest.raises(K.KDEError):
        K.resample(p, 10, ksType="discrete")


@pytest.mark.parametrize("d,M,N", [(2, 2, 100), (1, 2, 300), (3, 3, 64), (4, 2, 512), (2, 2, 2), (3, 2, 513), (2, 3, 1500)])
def test_product_in_one_call_equals_the_two_step_route(d, M, N):
    """kdeb200_product_kde (Gibbs kernel + on-chip sort / ball-tree statistics / golden sections on device-resident
    samples) against prodAppxMSGibbsS followed by kde!(pGM): same points, bit-identical bandwidths (N <= 512: fused
    kernel; above: the two-step route inside the call)."""
    rng = np.random.default_rng(d * 10 + N)
    trees = [K.kde(mixture(rng, d, N, 0.2 * j) if N > 3 else rng.normal(size=(d, N)), np.full(d, 0.4)) for j in range(M)]
    pq = K.prod(trees, seed=77)
    pts, _ = K.prodAppxMSGibbsS(None, trees, None, None, Niter=5, Np=N, seed=77)
    ref = K.kde(pts)
    assert np.array_equal(K.getPoints(pq), pts)
    assert np.array_equal(K.getBW(pq)[:, 0], K.getBW(ref)[:, 0]), (K.getBW(pq)[:, 0], K.getBW(ref)[:, 0])
    o = OKDE.kde_lcv(pts)
    assert relerr(K.getBW(pq)[:, 0] ** 2, o.arrays()["bandwidthMin"][:d]) < 1e-10


def test_kld_direct_and_unscented():
    """kld (src/DualTree01.jl:477-503): both methods, against the oracle assembled the same way."""
    rng = np.random.default_rng(21)
    a, b = rng.standard_normal((2, 120)), 0.4 + 1.2 * rng.standard_normal((2, 150))
    p, q = K.kde(a, [0.4, 0.5]), K.kde(b, [0.5, 0.4])
    op, oq = OKDE.kde_bw(a, [0.4, 0.5]), OKDE.kde_bw(b, [0.5, 0.4])
    exp = op.eval_avg_logl(op) - oq.eval_avg_logl(op)
    assert abs(K.kld(p, q) - exp) <= 1e-11 * abs(exp)
    D, N = 2, 120
    ptsE = np.tile(a, (1, 2 * D + 1))
    bw = np.array([[0.4], [0.5]]) * np.ones((2, N))
    for i in range(1, D + 1):
        ptsE[i - 1, (i - 1) * N:(i - 1) * N + N] += bw[i - 1]
        ptsE[i - 1, (2 * i - 1) * N:(2 * i - 1) * N + N] -= bw[i - 1]
    oE = OKDE.kde_lcv(ptsE)
    exp_u = op.eval_avg_logl(oE) - oq.eval_avg_logl(oE)
    got_u = K.kld(p, q, method="unscented")
    assert abs(got_u - exp_u) <= 1e-9 * abs(exp_u) and got_u > 0
    with pytest.raises(K.KDEError):
        K.kld(p, q, method="nope")
