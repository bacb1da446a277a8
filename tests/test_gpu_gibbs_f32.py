"""-m gpu: K1f, the FP32 variant of the Gibbs sampler (csrc/gibbs_f32.cu) -- mode (b) of BASELINE.json only.
Label probabilities are evaluated in packed FP32, so the bar is statistical: (1) fed the SAME Philox streams, a chain
follows the FP64 kernel until a uniform lands within ~1e-6 of a CDF edge, so nearly all samples carry identical labels
and (state, samplePoint! and output stay FP64) the same point to 1e-9; (2) against the oracle on independent streams
the product passes the two-sample KS / MMD tests and the reference's own statistical bands (test/runtests.jl:167-201)."""
import os

import numpy as np
import pytest

import kde_b200 as K
from oracle import oracle as O
from tests import test_oracle_stats as STATS
from tests.util import mixture, silverman

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def thread_per_chain_kernel():
    """K1f replaces the thread-per-chain kernel only; small calls would otherwise go to the (FP64) warp-per-chain kernel."""
    old = os.environ.get("KDEB200_GIBBS_WARP_MAX")
    os.environ["KDEB200_GIBBS_WARP_MAX"] = "0"
    K.gibbs_f32_slow_draws()
    yield
    K.set_gibbs_precision(K.F64)
    if old is None:
        os.environ.pop("KDEB200_GIBBS_WARP_MAX", None)
    else:
        os.environ["KDEB200_GIBBS_WARP_MAX"] = old


def both(trees, Np, T, seed, **kw):
    K.set_gibbs_precision(K.F64)
    p64, i64 = K.prodAppxMSGibbsS(None, trees, None, None, Niter=T, Np=Np, seed=seed, **kw)
    K.set_gibbs_precision(K.F32)
    p32, i32 = K.prodAppxMSGibbsS(None, trees, None, None, Niter=T, Np=Np, seed=seed, **kw)
    K.set_gibbs_precision(K.F64)
    return p64, i64, p32, i32


def agreement(p64, i64, p32, i32):
    same = np.all(i64 == i32, axis=0)
    if same.any():
        scale = np.maximum(np.abs(p64[:, same]), 1e-3 * np.max(np.abs(p64)) + 1e-300)
        assert float(np.max(np.abs(p32[:, same] - p64[:, same]) / scale)) < 1e-9  # same labels, same normals => same point
    return float(same.mean())


@pytest.mark.parametrize("d,M,N,Np,T", [
    (3, 8, 4096, 4096, 5),     # the C4 shape
    (1, 2, 300, 3000, 5), (2, 2, 100, 2000, 5), (3, 6, 100, 1500, 5), (4, 3, 500, 1000, 3),
    (5, 3, 200, 600, 2), (6, 2, 300, 500, 2), (7, 2, 128, 400, 1), (8, 2, 150, 400, 2),
    (2, 3, 1, 300, 3), (3, 1, 50, 300, 3), (2, 16, 20, 200, 1), (2, 2, 5001, 300, 1), (1, 3, 2, 300, 5),
    (2, 2, 100_000, 200, 1),   # checkpoint chunks of 2048 nodes spanning several tiles, pass 2 over long chunks
])
def test_same_streams_same_chains_as_fp64(d, M, N, Np, T):
    rng = np.random.default_rng(100 * d + M + N)
    trees = []
    for j in range(M):
        p = mixture(rng, d, N, 0.25 * j)
        trees.append(K.kde(p, silverman(p) if N > 2 else np.full(d, 0.7)))
    frac = agreement(*both(trees, Np, T, seed=7 + d))
    assert frac >= 0.97, frac
    assert K.gibbs_f32_slow_draws() == 0  # ordinary products never need the FP64 redo


def test_weights_masks_and_no_entropy():
    rng = np.random.default_rng(5)
    d, N = 3, 200
    trees = []
    for j in range(4):
        p = mixture(rng, d, N, 0.25 * j)
        w = rng.random(N) + 0.1
        trees.append(K.kde(p, silverman(p), w / w.sum()))
    mask = [[True, True, False], [True, True, True], [False, True, True], [True, False, True]]
    assert agreement(*both(trees, 1500, 4, seed=3, partialDimMask=mask)) >= 0.97
    assert agreement(*both(trees, 1500, 4, seed=4, addEntropy=False)) >= 0.97


def test_poisoned_masked_dimensions_do_not_leak():
    """test/testPartialProd.jl's construction: masked-out coordinates hold 9999999.0; the FP32 affine map ignores them."""
    rng = np.random.default_rng(99)
    (pts1, pts2, pts3), mask = STATS.partial_prod_case(rng)
    bw1, bw3 = K.getBW(K.kde(pts1))[:, 0], K.getBW(K.kde(pts3))[:, 0]
    P2 = K.kde(pts2)
    pts1[1, :] = 9999999.0
    pts3[0, :] = 9999999.0
    P = [K.kde(pts1, bw1), P2, K.kde(pts3, bw3)]
    p64, i64, p32, i32 = both(P, 2000, 3, seed=5, partialDimMask=mask)
    assert agreement(p64, i64, p32, i32) >= 0.97
    assert 0.8 * 2000 < np.sum((0 < p32[0, :]) & (p32[0, :] < 10))
    assert 0.8 * 2000 < np.sum((-10 < p32[1, :]) & (p32[1, :] < 0))


def test_fp32_underflow_is_redone_in_fp64():
    """Two densities 40 bandwidths apart: every leaf-level FP32 total underflows, the lanes redo those draws with the
    reference's literal FP64 arithmetic and the result still follows the FP64 kernel."""
    rng = np.random.default_rng(8)
    a = rng.standard_normal((2, 300)) * 0.05
    b = rng.standard_normal((2, 300)) * 0.05 + np.array([[3.0], [0.0]])
    trees = [K.kde(a, [0.02]), K.kde(b, [0.02])]
    p64, i64, p32, i32 = both(trees, 600, 3, seed=11)
    slow = K.gibbs_f32_slow_draws()
    assert slow > 0
    assert agreement(p64, i64, p32, i32) >= 0.9


def test_refuses_bandwidths_fp32_cannot_normalise():
    rng = np.random.default_rng(9)
    p = rng.standard_normal((2, 100))
    trees = [K.kde(p, [1e-6]), K.kde(p + 0.1, [0.3])]
    K.set_gibbs_precision(K.F32)
    with pytest.raises(K.KDEError, match="FP32"):
        K.prodAppxMSGibbsS(None, trees, None, None, Niter=2, Np=500, seed=1)
    K.set_gibbs_precision(K.F64)
    K.prodAppxMSGibbsS(None, trees, None, None, Niter=2, Np=500, seed=1)  # the FP64 sampler takes it


@pytest.mark.parametrize("d,M,N", [(1, 2, 150), (2, 3, 200), (3, 4, 128)])
def test_free_running_ks_and_mmd_against_the_oracle(d, M, N):
    from scipy.stats import ks_2samp
    from tests.test_gpu_gibbs import _mmd2_unbiased
    rng = np.random.default_rng(777 + d)
    kt, ot = [], []
    for j in range(M):
        p = mixture(rng, d, N, 0.25 * j)
        kt.append(K.kde(p, silverman(p)))
        ot.append(O.OKDE.kde_bw(p, silverman(p)))
    Np, T = 1500, 5
    K.set_gibbs_precision(K.F32)
    gp, _ = K.prodAppxMSGibbsS(None, kt, None, None, Niter=T, Np=Np, seed=99)
    K.set_gibbs_precision(K.F64)
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


@pytest.mark.parametrize("case", STATS.RANGE_UNIT_TESTS[:4], ids=lambda c: "-".join("%s%d" % kv for kv in c.items()))
def test_reference_range_unit_tests_fp32(case):
    """rangeUnitTests (test/runtests.jl:184-201) at the reference's own 5-of-10 criterion through the FP32 sampler."""
    rng = np.random.default_rng(1234 + 17 * case["D"] + case["M"])

    def prod(P, n, MCMC, rng):
        dummy = K.kde(rng.standard_normal((K.Ndim(P[0]), n)), [1.0])
        K.set_gibbs_precision(K.F32)
        try:
            return K.prodAppxMSGibbsS(dummy, P, None, None, Niter=MCMC, seed=int(rng.integers(1, 2 ** 62)))[0]
        finally:
            K.set_gibbs_precision(K.F64)
    v = [STATS.test_prods(prod, K.kde, rng, **case) for _ in range(10)]
    assert sum(v) >= 5, v
