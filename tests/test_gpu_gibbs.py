"""-m gpu: K1 parity.  Mode (a) of BASELINE.json: identical injected uniform/normal variates =>
labels identical to the oracle, product points within 1e-10 relative (FP64).  Mode (b): the
free-running Philox stream is reproduced exactly by injecting kdeb200_philox_streams() into the
oracle, plus the reference's own statistical bands (test/runtests.jl:167-201)."""
import numpy as np
import pytest

import kde_b200 as K
from oracle import oracle as O
from tests import test_oracle_stats as STATS
from tests.util import mixture, silverman

pytestmark = pytest.mark.gpu

# Both Gibbs kernels must pass everything: "thread" = one thread per chain (K1, the throughput kernel), "warp" = one warp
# per chain (K1w, the small-call kernel).  The library picks by the number of chains; KDEB200_GIBBS_WARP_MAX overrides.
HEAVY = ("test_full_size_c4_properties", "test_very_large_trees_multi_tile_chunks", "test_large_trees_c4_shape")


@pytest.fixture(autouse=True, params=["thread", "warp"])
def gibbs_kernel_choice(request):
    import os
    if request.param == "warp" and request.node.name.split("[")[0] in HEAVY:
        pytest.skip("sized for the thread-per-chain kernel (level lists beyond the warp kernel's shared memory, or 1M chains)")
    old = os.environ.get("KDEB200_GIBBS_WARP_MAX")
    os.environ["KDEB200_GIBBS_WARP_MAX"] = "0" if request.param == "thread" else "1000000000"
    yield request.param
    if old is None:
        os.environ.pop("KDEB200_GIBBS_WARP_MAX", None)
    else:
        os.environ["KDEB200_GIBBS_WARP_MAX"] = old
PT_TOL = 1e-10


def make(rng, d, N, shift=0.0, bw=None, weights=None):
    pts = mixture(rng, d, N, shift)
    if bw is None:
        bw = silverman(pts) if N > 2 else np.full(d, 0.7)
    return K.kde(pts, bw, weights), O.OKDE.kde_bw(pts, bw, weights)


def compare(ktrees, otrees, Np, T, rng, add_entropy=True, mask=None):
    nU, nN = O.prod_sizes(otrees, Np, T)
    U, G = rng.random(nU), rng.standard_normal(nN)
    epts, eind = O.gibbs(otrees, Np, T, U, G, add_entropy=add_entropy, mask=mask)
    dummy = K.kde(np.zeros((ktrees[0].bt.dims, Np)) + np.arange(Np), [1.0]) if Np > 0 else None
    gpts, gind = K.prodAppxMSGibbsS(dummy, ktrees, None, None, Niter=T, randU=U, randN=G, addEntropy=add_entropy,
                                    partialDimMask=mask)
    assert gind.shape == eind.shape and gpts.shape == epts.shape
    nbad = int(np.sum(np.any(gind != eind, axis=0)))
    assert nbad == 0, "label mismatches in %d of %d samples" % (nbad, Np)
    scale = np.maximum(np.abs(epts), 1e-3 * np.max(np.abs(epts)) + 1e-300)
    assert float(np.max(np.abs(gpts - epts) / scale)) < PT_TOL
    return gpts, gind


@pytest.mark.parametrize("d,M,N,Np,T", [
    (2, 2, 100, 100, 5),      # README product (BASELINE config 1)
    (3, 6, 100, 100, 5),      # testProds default
    (2, 4, 100, 100, 5), (4, 6, 100, 200, 10), (3, 5, 300, 100, 5), (2, 7, 100, 300, 5), (3, 2, 100, 100, 100),
    (1, 2, 64, 150, 3), (3, 3, 33, 130, 5), (5, 3, 50, 64, 2), (8, 2, 40, 64, 2), (1, 3, 2, 50, 5),
    (6, 2, 700, 40, 2), (7, 3, 300, 40, 1), (4, 2, 2000, 64, 2), (2, 2, 5000, 40, 1),
    (2, 3, 1, 40, 3), (3, 1, 50, 40, 3), (2, 16, 20, 33, 1), (3, 2, 100, 1, 5), (2, 2, 100, 100, 0),
])
def test_injected_streams_match_oracle(d, M, N, Np, T):
    rng = np.random.default_rng(1000 * d + 10 * M + N + T)
    pairs = [make(rng, d, N, 0.25 * j) for j in range(M)]
    compare([p[0] for p in pairs], [p[1] for p in pairs], Np, T, rng)


def test_mixed_sizes_beta_rayleigh_shape():
    """BASELINE config 2 shape: 1-D, 300 and 100 components (level lists of different depth)."""
    rng = np.random.default_rng(2)
    a = rng.beta(1.0, 0.45, size=(1, 300))
    b = rng.rayleigh(0.5, size=(1, 100)) - 0.5
    ka, oa = K.kde(a, [0.05]), O.OKDE.kde_bw(a, [0.05])
    kb, ob = K.kde(b, [0.08]), O.OKDE.kde_bw(b, [0.08])
    compare([ka, kb], [oa, ob], 2000, 5, rng)


def test_large_trees_c4_shape():
    """BASELINE config 4 shape at a sample count the oracle finishes in seconds."""
    rng = np.random.default_rng(4)
    pairs = [make(rng, 3, 4096, 0.25 * j) for j in range(8)]
    compare([p[0] for p in pairs], [p[1] for p in pairs], 192, 5, rng)


def test_weighted_and_lcv_bandwidth():
    rng = np.random.default_rng(8)
    pts = [rng.standard_normal((2, 100)) + s for s in (0.0, 2.0)]
    kt = [K.kde(p) for p in pts]
    ot = [O.OKDE.kde_lcv(p) for p in pts]
    compare(kt, ot, 100, 5, rng)
    w = rng.random(100) + 0.01
    kt = [K.kde(p, [0.3, 0.4], w) for p in pts]
    ot = [O.OKDE.kde_bw(p, [0.3, 0.4], w) for p in pts]
    compare(kt, ot, 100, 5, rng)


def test_add_entropy_false_is_precision_weighted_mean():
    rng = np.random.default_rng(11)
    pairs = [make(rng, 2, 50, 0.5 * j) for j in range(3)]
    kt, ot = [p[0] for p in pairs], [p[1] for p in pairs]
    pts, ind = compare(kt, ot, 64, 5, rng, add_entropy=False)
    lam = np.stack([1.0 / (K.getBW(t)[:, 0] ** 2) for t in kt])          # M x d
    for s in range(64):
        mus = np.stack([K.getPoints(t)[:, ind[j, s] - 2] for j, t in enumerate(kt)])  # label = index + 1
        exp = (lam * mus).sum(0) / lam.sum(0)
        assert np.allclose(pts[:, s], exp, rtol=1e-12, atol=1e-14)


def test_partial_dim_mask():
    """test/testPartialProd.jl: masked dims poisoned with 9999999 must not leak."""
    rng = np.random.default_rng(12)
    p1, p2, p3 = rng.random((2, 100)) + 10.0, rng.random((2, 100)), rng.random((2, 100)) - 10.0
    bw = [0.1, 0.1]
    p1[1, :] = 9999999.0
    p3[0, :] = 9999999.0
    mask = [[True, False], [True, True], [False, True]]
    kt = [K.kde(p, bw) for p in (p1, p2, p3)]
    ot = [O.OKDE.kde_bw(p, bw) for p in (p1, p2, p3)]
    pts, _ = compare(kt, ot, 100, 3, rng, mask=mask)
    assert np.sum((0 < pts[0]) & (pts[0] < 10)) > 80 and np.sum((-10 < pts[1]) & (pts[1] < 0)) > 80


def test_two_single_kernels_give_gaussian_product():
    p1, p2 = K.kde(np.array([[1.0]]), [0.5]), K.kde(np.array([[3.0]]), [1.0])
    dummy = K.kde(np.zeros((1, 4000)) + np.arange(4000), [1.0])
    pts, ind = K.prodAppxMSGibbsS(dummy, [p1, p2], None, None, Niter=5, seed=7)
    v = 1 / (1 / .25 + 1)
    m = v * (1 / .25 + 3)
    assert abs(pts.mean() - m) < 0.03 and abs(pts.var() - v) < 0.02 and np.all(ind == 2)


def test_philox_mode_is_reproduced_by_injection_and_by_oracle():
    rng = np.random.default_rng(21)
    pairs = [make(rng, 3, 200, 0.25 * j) for j in range(4)]
    kt, ot = [p[0] for p in pairs], [p[1] for p in pairs]
    Np, T, seed = 300, 5, 20261017
    L, perU, perN, evals = K.gibbs_sizes(kt, T)
    assert L == O.gibbs_nlevels(ot) and perU == 4 * (1 + L * (1 + T)) and perN == 3 * (L + 1)
    dummy = K.kde(np.zeros((3, Np)) + np.arange(Np), [1.0])
    p0, i0 = K.prodAppxMSGibbsS(dummy, kt, None, None, Niter=T, seed=seed)
    U, G = K.philox_streams(seed, Np, perU, perN)
    assert 0.0 <= U.min() and U.max() < 1.0 and abs(U.mean() - 0.5) < 0.01 and abs(G.std() - 1.0) < 0.02
    p1, i1 = K.prodAppxMSGibbsS(dummy, kt, None, None, Niter=T, randU=U, randN=G)
    assert np.array_equal(i0, i1) and np.array_equal(p0, p1)
    ep, ei = O.gibbs(ot, Np, T, U, G)
    assert np.array_equal(i0, ei) and np.max(np.abs(p0 - ep)) < 1e-10 * np.max(np.abs(ep))
    # sharding invariance: any split of the sample range gives the same chains
    a, ai = K.prodAppxMSGibbsS(dummy, kt, None, None, Niter=T, seed=seed, s0=0, s1=77)
    b, bi = K.prodAppxMSGibbsS(dummy, kt, None, None, Niter=T, seed=seed, s0=77, s1=Np)
    assert np.array_equal(np.hstack([a, b]), p0) and np.array_equal(np.hstack([ai, bi]), i0)
    a, ai = K.prodAppxMSGibbsS(dummy, kt, None, None, Niter=T, randU=U, randN=G, s0=130, s1=Np)
    assert np.array_equal(a, p0[:, 130:]) and np.array_equal(ai, i0[:, 130:])


def _gpu_prod(P, n, MCMC, rng):
    dummy = K.kde(rng.standard_normal((K.Ndim(P[0]), n)), [1.0])
    return K.prodAppxMSGibbsS(dummy, P, None, None, Niter=MCMC, seed=int(rng.integers(1, 2 ** 62)))[0]


@pytest.mark.parametrize("case", STATS.RANGE_UNIT_TESTS, ids=lambda c: "-".join("%s%d" % kv for kv in c.items()))
def test_reference_range_unit_tests(case):
    """rangeUnitTests (test/runtests.jl:184-201) on the CUDA path at the reference's own criterion: every one of
    its eight shape sets, 10 repetitions of testProds each (LOOCV kde! of fresh normal data, free-running
    Philox), at least 5 of 10 inside the bands.  tests/test_oracle_stats.py runs the same on the oracle."""
    rng = np.random.default_rng(4321 + 17 * case["D"] + case["M"])
    v = [STATS.test_prods(_gpu_prod, K.kde, rng, **case) for _ in range(10)]
    assert sum(v) >= 5, v


def test_reference_partial_product_free_running():
    """test/testPartialProd.jl:8-58 on the CUDA path with its own defaults (Niter = 3, LOOCV bandwidths of the
    unpoisoned points, free-running RNG): > 80 of 100 samples inside (0, 10) x (-10, 0)."""
    rng = np.random.default_rng(99)
    (pts1, pts2, pts3), mask = STATS.partial_prod_case(rng)
    bw1, bw3 = K.getBW(K.kde(pts1))[:, 0], K.getBW(K.kde(pts3))[:, 0]
    P2 = K.kde(pts2)
    pts1[1, :] = 9999999.0
    pts3[0, :] = 9999999.0
    P = [K.kde(pts1, bw1), P2, K.kde(pts3, bw3)]
    dummy = K.kde(rng.random((2, 100)))
    pGM, _ = K.prodAppxMSGibbsS(dummy, P, None, None, partialDimMask=mask, seed=5)
    assert 80 < np.sum((0 < pGM[0, :]) & (pGM[0, :] < 10))
    assert 80 < np.sum((-10 < pGM[1, :]) & (pGM[1, :] < 0))


def test_sixteen_densities_simd_sum_edge():
    """M = 16 = KDEB200_MAX_DENS is where Julia's sum(lambdas) switches to an @simd loop (src/MSGibbs01.jl:141):
    the kernel sums sequentially like the default oracle; against the oracle's emulated 2x4-lane vector sum the
    labels still agree (points to 1e-12) -- the summation order moves cov by <= 1 ulp."""
    rng = np.random.default_rng(1616)
    pairs = [make(rng, 2, 24, 0.1 * j, bw=np.array([0.4 + 0.031 * j, 0.5 + 0.017 * j])) for j in range(16)]
    kt, ot = [p[0] for p in pairs], [p[1] for p in pairs]
    Np, T = 500, 2
    nU, nN = O.prod_sizes(ot, Np, T)
    U, G = rng.random(nU), rng.standard_normal(nN)
    gp, gi = K.prodAppxMSGibbsS(None, kt, None, None, Niter=T, Np=Np, randU=U, randN=G)
    try:
        O.set_sum_simd(2, 4)
        ep, ei = O.gibbs(ot, Np, T, U, G)
    finally:
        O.set_sum_simd(0, 0)
    bad = np.any(gi != ei, axis=0)
    assert bad.mean() <= 0.01
    assert np.max(np.abs(gp[:, ~bad] - ep[:, ~bad])) < 1e-10 * np.max(np.abs(ep))


def test_product_operator_and_errors():
    rng = np.random.default_rng(3)
    p, q = K.kde(rng.standard_normal((2, 100))), K.kde(2.0 + rng.standard_normal((2, 100)))
    pq = p * q
    assert K.Npts(pq) == 100 and K.Ndim(pq) == 2
    assert np.all(np.abs(K.getPoints(pq).mean(axis=1) - 1.0) < 0.5)
    with pytest.raises(K.KDEError):
        K.prodAppxMSGibbsS(p, [p, K.kde(rng.standard_normal((3, 10)), [1.0])], None, None)
    with pytest.raises(K.KDEError):
        K.prodAppxMSGibbsS(p, [p, q], None, None, randU=np.zeros(10), randN=np.zeros(10))
    with pytest.raises(K.KDEError):
        K.prodAppxMSGibbsS(p, [p, q], None, None, getMu=(lambda *a: 0,))


def test_label_recording_matches_oracle():
    """glbs.recordChoosen / labelsChoosen[sample][density][level] (src/MSGibbs01.jl:29-31,109-112)."""
    rng = np.random.default_rng(31)
    pairs = [make(rng, 2, 37, 0.3 * j) for j in range(3)]
    kt, ot = [p[0] for p in pairs], [p[1] for p in pairs]
    Np, T = 150, 4
    nU, nN = O.prod_sizes(ot, Np, T)
    U, G = rng.random(nU), rng.standard_normal(nN)
    ep, ei, er = O.gibbs(ot, Np, T, U, G, record=True)
    gp, gi, gr = K.prodAppxMSGibbsS(None, kt, None, None, Niter=T, Np=Np, randU=U, randN=G, recordLabels=True)
    assert gr.shape == er.shape == (Np, 3, O.gibbs_nlevels(ot))
    assert np.array_equal(gr, er) and np.array_equal(gi, ei)
    assert np.array_equal(gr[:, :, -1] + 1, gi.T)          # the last level's record is the output label - 1
    assert (gr[:, :, 0] == 0).all()                        # level 1 nodes are internal: permutation 0
    _, _, g0 = K.prodAppxMSGibbsS(None, kt, None, None, Niter=0, Np=8, seed=1, recordLabels=True)
    assert (g0 == -1).all()                                # sampleIndex never ran


def test_full_size_c4_properties():
    """BASELINE config 4 at full size (8 x 4096 components, 3-D, 1M samples, Niter=5): size-independent
    properties -- determinism, sharding invariance on slices, label range, agreement of the product
    moments between two independent seeds, and the exact point/label relation with addEntropy=false."""
    import bench
    trees = [K.kde(bench.synth_points(j), bench.silverman(bench.synth_points(j))) for j in range(bench.NDENS)]
    Np = 1_000_000
    p1, i1 = K.prodAppxMSGibbsS(None, trees, None, None, Niter=5, Np=Np, seed=11)
    assert p1.shape == (3, Np) and i1.shape == (8, Np) and np.isfinite(p1).all()
    assert i1.min() >= 2 and i1.max() <= 4097
    a, ai = K.prodAppxMSGibbsS(None, trees, None, None, Niter=5, Np=Np, seed=11, s0=123_456, s1=125_000)
    assert np.array_equal(a, p1[:, 123_456:125_000]) and np.array_equal(ai, i1[:, 123_456:125_000])
    p2, _ = K.prodAppxMSGibbsS(None, trees, None, None, Niter=5, Np=Np, seed=12)
    se = p1.std(axis=1) / np.sqrt(Np / 50.0)               # generous: chains are independent across samples
    assert np.all(np.abs(p1.mean(axis=1) - p2.mean(axis=1)) < 6 * se)
    assert np.all(np.abs(p1.std(axis=1) / p2.std(axis=1) - 1) < 0.02)
    q, qi = K.prodAppxMSGibbsS(None, trees, None, None, Niter=5, Np=4096, seed=11, addEntropy=False)
    lam = np.stack([1.0 / (K.getBW(t)[:, 0] ** 2) for t in trees])
    pts = [K.getPoints(t) for t in trees]
    for s in range(0, 4096, 97):
        mus = np.stack([pts[j][:, qi[j, s] - 2] for j in range(8)])
        assert np.allclose(q[:, s], (lam * mus).sum(0) / lam.sum(0), rtol=1e-12, atol=1e-14)


def test_extracting_labels_example():
    """examples/ExtractingLabels.jl: with addEntropy=false each product point is the mean of the three
    selected kernel means, recoverable from the recorded labels of the last level."""
    X = [K.kde(np.array([[1.0, 2, 3]]), [1.0]), K.kde(np.array([[0.5, 1.5, 2.5]]), [1.0]), K.kde(np.array([[4.0, 5, 6]]), [1.0])]
    # LOOCV of the 3 product points is degenerate; take the raw samples instead of kde!(pGM)
    pts, ind, lab = K.prodAppxMSGibbsS(None, X, None, None, Niter=5, Np=3, addEntropy=False, seed=3, recordLabels=True)
    for s in range(3):
        mu = np.mean([K.getPoints(X[j])[0, lab[s, j, -1] - 1] for j in range(3)])
        assert abs(pts[0, s] - mu) < 1e-13


@pytest.mark.parametrize("M,N,T,Np", [(1, 1, 0, 5000), (1, 2, 1, 4097), (2, 2, 0, 3000), (2, 3, 1, 2500), (3, 5, 0, 129)])
def test_many_batches_tiny_schedules(M, N, T, Np):
    """Dynamic batch scheduling with schedules of only a few tiles per batch (the tile ring's look-ahead
    crosses batch boundaries constantly) and sample counts that are not multiples of the CTA size."""
    rng = np.random.default_rng(M * 100 + N * 10 + T)
    pairs = [make(rng, 2, N, 0.5 * j, bw=np.array([0.5, 0.7])) for j in range(M)]
    compare([p[0] for p in pairs], [p[1] for p in pairs], Np, T, rng)


def _mmd2_unbiased(X, Y, gamma):
    """unbiased MMD^2 with an RBF kernel exp(-gamma |x-y|^2); X, Y are n x d"""
    def gram(A, B):
        d2 = (A * A).sum(1)[:, None] + (B * B).sum(1)[None, :] - 2.0 * A @ B.T
        return np.exp(-gamma * np.maximum(d2, 0.0))
    n, m = len(X), len(Y)
    Kxx, Kyy, Kxy = gram(X, X), gram(Y, Y), gram(X, Y)
    return ((Kxx.sum() - np.trace(Kxx)) / (n * (n - 1)) + (Kyy.sum() - np.trace(Kyy)) / (m * (m - 1))
            - 2.0 * Kxy.mean())


@pytest.mark.parametrize("d,M,N", [(1, 2, 150), (2, 3, 200), (3, 4, 128)])
def test_free_running_rng_two_sample_ks_and_mmd(d, M, N):
    """Mode (b) of BASELINE.json: with free-running RNG (device Philox vs numpy streams fed to the oracle) the
    two product sample sets pass a two-sample Kolmogorov-Smirnov test per dimension and a kernel MMD permutation
    test.  Independent streams, so the samples differ; only their law must agree."""
    from scipy.stats import ks_2samp
    rng = np.random.default_rng(4242 + d)
    pairs = [make(rng, d, N, 0.25 * j) for j in range(M)]
    kt, ot = [p[0] for p in pairs], [p[1] for p in pairs]
    Np, T = 1500, 5
    gp, _ = K.prodAppxMSGibbsS(None, kt, None, None, Niter=T, Np=Np, seed=99)
    nU, nN = O.prod_sizes(ot, Np, T)
    ep, _ = O.gibbs(ot, Np, T, rng.random(nU), rng.standard_normal(nN))
    for k in range(d):
        assert ks_2samp(gp[k], ep[k]).pvalue > 1e-3, "KS rejects equality of dimension %d" % k
    X, Y = gp.T, ep.T
    Z = np.vstack([X, Y])
    med = np.median(((Z[:400, None, :] - Z[None, :400, :]) ** 2).sum(-1))
    gamma = 1.0 / max(med, 1e-12)
    stat = _mmd2_unbiased(X, Y, gamma)
    null = []
    for _ in range(100):
        perm = rng.permutation(len(Z))
        null.append(_mmd2_unbiased(Z[perm[:Np]], Z[perm[Np:]], gamma))
    assert stat <= np.quantile(null, 0.99) + 3 * np.std(null), (stat, np.quantile(null, 0.99))
    # and a shifted law IS detected by the same statistic (the test has power)
    assert _mmd2_unbiased(X + 0.3 * X.std(0), Y, gamma) > np.quantile(null, 0.99) + 3 * np.std(null)


def test_very_large_trees_multi_tile_chunks():
    """100 000-component trees: levels of up to 1e5 nodes (hundreds of tiles per draw, checkpoint chunks of 2048 nodes
    spanning several tiles, pass 2 over long chunks) keep labels exact and points within 1e-10."""
    rng = np.random.default_rng(31337)
    pairs = [make(rng, 2, 100_000, 0.25 * j) for j in range(2)]
    compare([p[0] for p in pairs], [p[1] for p in pairs], 130, 1, rng)


DEGENERATE_BW = {
    "zero_in_one_dim": [[0.3, 0.3], [0.0, 0.3], [0.3, 0.3]],
    "zero_everywhere": [[0.0, 0.0], [0.3, 0.3], [0.3, 0.3]],
    "nan_in_one_dim": [[np.nan, 0.3], [0.3, 0.3], [0.3, 0.3]],
    "inf_in_one_dim": [[np.inf, 0.3], [0.3, 0.3], [0.3, 0.3]],
    "variance_1e50": [[1e25, 0.3], [0.3, 0.3], [0.3, 0.3]],
    "variance_1e-50": [[1e-25, 0.3], [0.3, 0.3], [1e-25, 1e-25]],
}


@pytest.mark.parametrize("case", sorted(DEGENERATE_BW))
def test_degenerate_bandwidths_follow_the_reference_nan_rules(case):
    """zero / NaN / Inf / out-of-range bandwidths used to be refused (code 8); the reference skips NaN
    terms, zeroes NaN likelihoods and falls back to the last node's weight (src/MSGibbs01.jl:287-315).  The library now runs
    such products through the verbatim-arithmetic variant of the warp-per-chain kernel (whatever the chain count): labels
    identical to the oracle, points identical including where they are NaN."""
    rng = np.random.default_rng(77)
    d, N, Np, T = 2, 60, 150, 3
    kt, ot = [], []
    for j, bw in enumerate(DEGENERATE_BW[case]):
        pts = mixture(rng, d, N, 0.25 * j)
        kt.append(K.kde(pts, bw))
        ot.append(O.OKDE.kde_bw(pts, bw))
    nU, nN = O.prod_sizes(ot, Np, T)
    U, G = rng.random(nU), rng.standard_normal(nN)
    with np.errstate(all="ignore"):
        epts, eind = O.gibbs(ot, Np, T, U, G)
    gpts, gind = K.prodAppxMSGibbsS(None, kt, None, None, Niter=T, Np=Np, randU=U, randN=G)
    assert np.array_equal(gind, eind)
    assert np.array_equal(np.isnan(gpts), np.isnan(epts)) and np.array_equal(np.isinf(gpts), np.isinf(epts))
    fin = np.isfinite(epts)
    if fin.any():
        assert np.max(np.abs(gpts[fin] - epts[fin]) / np.maximum(np.abs(epts[fin]), 1e-3)) < PT_TOL
    if case.startswith("variance"):
        assert fin.all()  # out-of-range but finite: ordinary results, only the arithmetic route differs
