"""CPU (not gpu): the reference's own STATISTICAL suite for the Gibbs product sampler, run on the oracle.

The reference holds no label / point goldens for prodAppxMSGibbsS -- its tests are statistical:
test/runtests.jl:167-201 (testProds / rangeTestProds / rangeUnitTests: eight shape sets, 10 repetitions each,
at least 5 must pass) and test/testPartialProd.jl:8-58 (partial-dimension masks, > 80 of 100 samples inside the
band).  Passing them at the reference's own criterion is the strongest pin the Gibbs half of the oracle can get
without a Julia run (julia/make_goldens.jl turns "unpinned" into "pinned" when one is available).  The same
suite runs on the GPU path in tests/test_gpu_gibbs.py.

Also here: the one known fidelity edge, Julia's @simd sum(lambdas) for Ndens >= 16 (src/MSGibbs01.jl:141).
"""
import numpy as np
import pytest

from oracle import oracle as O

# rangeUnitTests (test/runtests.jl:189-201), the commented-out (D=3, M=10) line left out as in the reference
RANGE_UNIT_TESTS = [
    dict(D=2, M=2), dict(D=2, M=4), dict(D=2, M=6), dict(D=3, M=6, MCMC=10), dict(D=4, M=6, n=200, MCMC=10),
    dict(D=3, M=5, N=300), dict(D=2, M=7, n=300), dict(D=3, M=2, MCMC=100),
]


def test_prods(prod_fn, kde_fn, rng, D=3, M=6, N=100, n=100, dev=1.0, MCMC=5):
    """testProds (test/runtests.jl:167-182) with the sampler and kde! injected (oracle here, CUDA path in the
    gpu suite): |mean| < prodDev and every per-dimension std within (0.66, 1.33) prodDev."""
    P = [kde_fn(dev * rng.standard_normal((D, N))) for _ in range(M)]
    pGM = prod_fn(P, n, MCMC, rng)
    if np.sum(np.abs(pGM)) < 1e-14:
        raise AssertionError("testProds -- prodAppxMSGibbsS, nothing in pGM")
    prodDev = np.sqrt(dev ** (2 * M) / (M * dev ** 2))
    T1 = np.linalg.norm(pGM.mean(axis=1)) < 1.0 * prodDev
    T2 = all(0.66 * prodDev < np.std(pGM[i, :], ddof=1) < 1.33 * prodDev for i in range(D))
    return bool(T1 and T2)


test_prods.__test__ = False  # helper shared with the gpu suite, not a test by itself


def oracle_prod(P, n, MCMC, rng):
    nU, nN = O.prod_sizes(P, n, MCMC)
    return O.gibbs(P, n, MCMC, rng.random(nU), rng.standard_normal(nN), nthreads=O.max_threads())[0]


@pytest.mark.parametrize("case", RANGE_UNIT_TESTS, ids=lambda c: "-".join("%s%d" % kv for kv in c.items()))
def test_oracle_passes_reference_range_unit_tests(case):
    """rangeTestProds (test/runtests.jl:184-187): 10 repetitions, at least 5 pass."""
    rng = np.random.default_rng(1234 + 17 * case["D"] + case["M"])
    v = [test_prods(oracle_prod, O.OKDE.kde_lcv, rng, **case) for _ in range(10)]
    assert sum(v) >= 5, v


def partial_prod_case(rng):
    """test/testPartialProd.jl:8-44: three 2-D densities at +10 / 0 / -10, P1 blind on dim 2, P3 blind on dim 1,
    the blind coordinates poisoned with 9999999; bandwidths from LOOCV of the unpoisoned points."""
    pts1, pts2, pts3 = rng.random((2, 100)) + 10.0, rng.random((2, 100)), rng.random((2, 100)) - 10.0
    mask = [[True, False], [True, True], [False, True]]
    return (pts1, pts2, pts3), mask


def test_oracle_passes_reference_partial_product():
    """test/testPartialProd.jl:47-53: > 80 of 100 product samples inside (0,10) x (-10,0); default Niter = 3."""
    rng = np.random.default_rng(99)
    (pts1, pts2, pts3), mask = partial_prod_case(rng)
    bw1, bw3 = O.OKDE.kde_lcv(pts1).get_bw()[:, 0], O.OKDE.kde_lcv(pts3).get_bw()[:, 0]
    P2 = O.OKDE.kde_lcv(pts2)
    pts1[1, :] = 9999999.0
    pts3[0, :] = 9999999.0
    P = [O.OKDE.kde_bw(pts1, bw1), P2, O.OKDE.kde_bw(pts3, bw3)]
    nU, nN = O.prod_sizes(P, 100, 3)
    pGM, _ = O.gibbs(P, 100, 3, rng.random(nU), rng.standard_normal(nN), mask=mask)
    assert 80 < np.sum((0 < pGM[0, :]) & (pGM[0, :] < 10))
    assert 80 < np.sum((-10 < pGM[1, :]) & (pGM[1, :] < 0))


def test_simd_sum_edge_at_16_densities():
    """Julia sums the 16 lambdas of a 16-density product with an @simd loop that LLVM may reassociate
    (oracle/kde_oracle.c julia_sum).  KDEB200_MAX_DENS = 16 sits exactly on that edge.  Measure how often the
    summation order changes a label: sequential (what the library and the default oracle do, and what AVX2 /
    AVX-512 hosts execute for n = 16 because 14 remaining elements do not fill one VF*IC = 16 / 32 wide trip) vs
    one 2x4-lane vector trip (SSE2-class hosts).  Labels must agree on all but a vanishing fraction of draws,
    and product points to 1e-12 wherever the labels agree; M < 16 is bit-identical by construction.
    Measured: 0 label mismatches in 20 000 chains (1.9e6 draws), points differ by <= 1 ulp (4.4e-16)."""
    rng = np.random.default_rng(16)
    for M, expect_identical in ((15, True), (16, False)):
        P = [O.OKDE.kde_bw(rng.standard_normal((2, 20)) + 0.1 * j, [0.4 + 0.031 * j, 0.5 + 0.017 * j]) for j in range(M)]
        Np, T = 2000, 2
        nU, nN = O.prod_sizes(P, Np, T)
        U, G = rng.random(nU), rng.standard_normal(nN)
        try:
            O.set_sum_simd(0, 0)
            p0, i0 = O.gibbs(P, Np, T, U, G, nthreads=O.max_threads())
            O.set_sum_simd(2, 4)
            p1, i1 = O.gibbs(P, Np, T, U, G, nthreads=O.max_threads())
        finally:
            O.set_sum_simd(0, 0)
        bad = np.any(i0 != i1, axis=0)
        if expect_identical:
            assert not bad.any() and np.array_equal(p0, p1)
        else:
            assert not np.array_equal(p0, p1)  # the emulation is live: some point differs in its last bits
            assert bad.mean() <= 0.01, "summation order flipped labels in %d of %d chains" % (bad.sum(), Np)
            ok = ~bad
            assert np.max(np.abs(p0[:, ok] - p1[:, ok])) <= 1e-12 * np.max(np.abs(p0))
