"""-m gpu: K2 parity (evaluation, LOO evaluation, LOO likelihood, LOOCV) through the C-ABI.

FP64 tolerance (BASELINE.json north_star): 1e-12 relative against the oracle."""
import json
import math
import os

import numpy as np
import pytest

import kde_b200 as K
from oracle.oracle import OKDE
from tests.util import mixture, relerr, silverman

pytestmark = pytest.mark.gpu
TOL = 1e-12
FIX = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_fixtures.json")))


@pytest.mark.parametrize("d,N,M", [(1, 300, 500), (2, 100, 100), (3, 1000, 777), (4, 513, 64), (5, 200, 129),
                                   (8, 150, 257), (3, 5000, 3), (1, 1, 10), (2, 7, 1)])
def test_eval_matches_oracle(d, N, M):
    rng = np.random.default_rng(100 + d * 7 + N)
    pts = mixture(rng, d, N)
    bw = silverman(pts) if N > 2 else np.full(d, 0.5)
    w = rng.random(N) + 0.05
    pos = mixture(rng, d, M) * 1.2
    p = K.kde(pts, bw, w)
    o = OKDE.kde_bw(pts, bw, w)
    got = K.evaluateDualTree(p, pos)
    exp = o.evaluate(pos)
    assert relerr(got, exp) < TOL
    assert relerr(p(pos), exp) < TOL  # functor form


@pytest.mark.parametrize("d", [1, 2, 3])
def test_single_kernel_is_normal_pdf(d):
    mu = np.arange(1, d + 1, dtype=np.float64).reshape(d, 1)
    h = np.linspace(0.3, 0.9, d)
    p = K.kde(mu, h)
    x = mu + np.linspace(-3, 3, 25)[None, :] * h[:, None]
    exp = np.prod(np.exp(-0.5 * ((x - mu) / h[:, None]) ** 2) / (np.sqrt(2 * np.pi) * h[:, None]), axis=0)
    assert relerr(K.evaluateDualTree(p, x), exp) < 1e-13


def test_far_tail_and_underflow():
    p = K.kde(np.array([[0.0, 1.0]]), [0.1])
    o = OKDE.kde_bw(np.array([[0.0, 1.0]]), [0.1])
    x = np.array([[3.0, 3.7, 3.8, 5.0, 50.0]])  # exponents -450 ... beyond underflow
    got, exp = K.evaluateDualTree(p, x), o.evaluate(x)
    assert np.all((exp == 0) == (got == 0))
    nz = exp > 1e-290
    assert relerr(got[nz], exp[nz]) < 1e-11


@pytest.mark.parametrize("d,N", [(1, 100), (2, 300), (3, 1000), (1, 2500)])
def test_loo_eval_and_entropy(d, N):
    rng = np.random.default_rng(5 + d + N)
    pts = mixture(rng, d, N)
    bw = silverman(pts)
    w = rng.random(N) + 0.1
    p, o = K.kde(pts, bw, w), OKDE.kde_bw(pts, bw, w)
    assert relerr(K.evaluateDualTree(p, p), o.evaluate()) < TOL
    assert relerr(K.evaluateDualTree(p, pts, True), o.evaluate()) < TOL
    assert abs(K.entropy(p) - o.entropy()) <= TOL * abs(o.entropy())
    assert abs(K.evalAvgLogL(p, p) + o.entropy()) <= TOL * abs(o.entropy())
    q, oq = K.kde(pts[:, : N // 2] + 0.1, bw), OKDE.kde_bw(pts[:, : N // 2] + 0.1, bw)
    assert abs(K.evalAvgLogL(p, q) - o.eval_avg_logl(oq)) <= 1e-11 * abs(o.eval_avg_logl(oq))
    for a in (0.5, 1.0, 1.7):
        assert abs(K.nLOO_LL(a, p) - o.nloo_ll(a)) <= TOL * abs(o.entropy())


def test_zero_rule_gives_inf():
    p = K.kde(np.array([[0.0, 1000.0]]), [0.01])
    assert K.entropy(p) == math.inf and K.evalAvgLogL(p, p) == -math.inf


def test_lcv_golden_fixture():
    """kde!(x) on the reference's 100 fixed points: LOOCV variance 0.00272597 (test/runtests.jl:104-116)."""
    case = FIX["UnitTest1Dlcv01"]
    pts = np.array(case["points"])
    cnt = []
    p = K.kde(pts)
    o = OKDE.kde_lcv(pts)
    assert abs(p.bandwidthMin[0] - case["expected"]["bwMin"][0]) < 5e-9
    assert abs(p.bandwidthMin[0] - o.arrays()["bandwidthMin"][0]) < 1e-12
    exp = case["expected"]
    for mine, name in [(p.bt.centers, "centers"), (p.bt.ranges, "ranges"), (p.means, "means"), (p.bandwidth, "bandwidth")]:
        assert np.linalg.norm(mine - np.array(exp[name])) <= case["tol"]
    pp = K.ksize(K.marginal(K.kde(pts, [1.0]), [1]), _count=cnt)
    assert cnt == [20]


def test_lcv_multidim_matches_oracle():
    rng = np.random.default_rng(9)
    pts = mixture(rng, 3, 400)
    p, o = K.kde(pts), OKDE.kde_lcv(pts)
    assert relerr(p.bandwidthMin[:3], o.arrays()["bandwidthMin"][:3]) < 1e-10


def test_errors():
    p = K.kde(np.zeros((2, 5)) + np.arange(5), [1.0])
    with pytest.raises(K.KDEError):
        K.evaluateDualTree(p, np.zeros((3, 4)))
    with pytest.raises(K.KDEError):
        K.evaluateDualTree(p, np.zeros((2, 4)), addop=(lambda a, b: a + b,))
    with pytest.raises(K.KDEError):
        K.kde(np.zeros((9, 5)) + np.arange(5), [1.0])._dev()


@pytest.mark.parametrize("d,N,M", [(1, 300, 500), (2, 1000, 333), (3, 4096, 1000), (4, 513, 64), (6, 200, 129)])
def test_fp32_variant_within_1e5(d, N, M):
    """FP32 (MUFU ex2) variant: 1e-5 relative (BASELINE.json north_star)."""
    rng = np.random.default_rng(300 + d + N)
    pts = mixture(rng, d, N)
    bw = silverman(pts)
    pos = mixture(rng, d, M)
    p, o = K.kde(pts, bw), OKDE.kde_bw(pts, bw)
    exp = o.evaluate(pos)
    got = K.evaluateDualTree(p, pos, precision=K.F32)
    assert relerr(got, exp) < 1e-5
    assert relerr(K.evaluateDualTree(p, p, precision=K.F32), o.evaluate()) < 1e-5


def test_full_size_c5_spot_check():
    """BASELINE config 5 at full size (1M components, 3-D): 256 of the query points against the oracle
    (1e-12), FP32 variant within 1e-5, plus linearity in the weights (a size-independent property)."""
    rng = np.random.default_rng(20261017)
    N = 1_000_000
    pts = mixture(rng, 3, N)
    bw = silverman(pts)
    p = K.kde(pts, bw)
    pos = mixture(rng, 3, 50_000)
    got = K.evaluateDualTree(p, pos)
    o = OKDE.kde_bw(pts, bw)
    exp = o.evaluate(pos[:, :256], nthreads=8)
    assert relerr(got[:256], exp) < TOL
    assert relerr(K.evaluateDualTree(p, pos, precision=K.F32), got) < 1e-5
    w = rng.random(N) + 0.5
    pa, pb = K.kde(pts[:, : N // 2], bw, w[: N // 2]), K.kde(pts[:, N // 2:], bw, w[N // 2:])
    pw = K.kde(pts, bw, w)
    fa = w[: N // 2].sum() / w.sum()
    mix = fa * K.evaluateDualTree(pa, pos[:, :2000]) + (1 - fa) * K.evaluateDualTree(pb, pos[:, :2000])
    assert relerr(K.evaluateDualTree(pw, pos[:, :2000]), mix) < 1e-11


def test_host_wrappers_over_the_eval_seam():
    """integralAppxUnitTests (test/runtests.jl:203-223) through the GPU path, plus getKDEMax / kld."""
    rng = np.random.default_rng(17)

    def offs(o, dim=1, N=201):
        p = K.kde(rng.standard_normal((dim, 100)))
        pts = rng.standard_normal((dim, 150))
        pts[0, :] += o
        return K.intersIntgAppxIS(p, K.kde(pts), N=N)

    assert 0.2 < offs(0.0) < 0.35
    assert 0.1 < offs(1.0, N=1000) < 0.3
    assert 0.01 < offs(-2.0, N=1000) < 0.17
    assert 0.05 < offs(0.0, dim=2, N=101) < 0.15
    p = K.kde(3.0 + 0.5 * rng.standard_normal((2, 400)))
    assert np.all(np.abs(K.getKDEMax(p) - 3.0) < 0.5)
    q = K.kde(3.5 + 0.5 * rng.standard_normal((2, 400)))
    o_p, o_q = OKDE.kde_bw(K.getPoints(p), K.getBW(p)[:, 0]), OKDE.kde_bw(K.getPoints(q), K.getBW(q)[:, 0])
    exp = o_p.eval_avg_logl(o_p) - o_q.eval_avg_logl(o_p)
    assert abs(K.kld(p, q) - exp) < 1e-10 * abs(exp) and K.minkld(p, q) > 0


def test_eval_only_handle_then_gibbs_upgrade():
    """kdeb200_tree_create_eval carries the leaf records alone: evaluation and LOOCV work on it, a Gibbs call is
    refused (code 7), and the Python mirror upgrades the handle on its first Gibbs use with unchanged answers."""
    import ctypes as C
    from kde_b200 import _lib
    rng = np.random.default_rng(77)
    pts = mixture(rng, 3, 700)
    p = K.kde(pts, [0.3, 0.2, 0.4])
    q = rng.normal(size=(3, 50))
    before = K.evaluateDualTree(p, q)
    H0 = K.entropy(p)
    assert not p._handle_gibbs
    nb_eval = C.c_int64(0)
    _lib.check(_lib.lib().kdeb200_tree_info(p._dev(), None, None, None, C.byref(nb_eval)))
    arr = (_lib.tree_t * 1)(p._dev())
    L = C.c_int(0)
    rc = _lib.lib().kdeb200_gibbs_sizes(arr, 1, 3, C.byref(L), None, None, None)
    assert rc == 7 and b"tree_create_eval" in _lib.lib().kdeb200_last_error()
    pts_out, _ = K.prodAppxMSGibbsS(None, [p, p], None, None, Niter=2, Np=64, seed=5)[:2]
    assert p._handle_gibbs and pts_out.shape == (3, 64)
    nb_full = C.c_int64(0)
    _lib.check(_lib.lib().kdeb200_tree_info(p._dev(), None, None, None, C.byref(nb_full)))
    assert nb_full.value > 2 * nb_eval.value
    assert np.array_equal(before, K.evaluateDualTree(p, q)) and H0 == K.entropy(p)


@pytest.mark.parametrize("d,N", [(1, 2), (1, 3), (1, 100), (2, 100), (3, 257), (2, 512), (1, 513), (4, 700), (2, 3000)])
def test_native_lcv_equals_stepwise_mirror(d, N):
    """kdeb200_kde_lcv (fused single-launch golden section for N <= 512, host loop above) gives the bits of the
    call-by-call mirror of kde!(points) (src/KDE01.jl:13-23) and the oracle's bandwidths to 1e-10."""
    rng = np.random.default_rng(1000 * d + N)
    pts = mixture(rng, d, N) if N > 3 else rng.normal(size=(d, N))
    calls_native, calls_mirror = [], []
    bw = K.lcv_bandwidths(pts, _count=calls_native)
    p0 = K.kde(pts, [1.0])
    ref = np.zeros(d)
    for i in range(d):
        ref[i] = K.getBW(K.ksize(K.marginal(p0, [i + 1]), _count=calls_mirror))[0, 0]
    assert np.array_equal(bw, ref), (bw, ref)
    assert calls_native == calls_mirror
    o = OKDE.kde_lcv(pts)
    assert relerr(bw ** 2, o.arrays()["bandwidthMin"][:d]) < 1e-10
    assert np.array_equal(K.getBW(K.kde(pts))[:, 0], np.sqrt(bw ** 2))


def test_native_lcv_degenerate_inputs():
    """identical points: every LOO likelihood is +Inf-free but the bracket collapses to the 1e-6 floor; one point is
    refused like the reference (minimum over an empty range)."""
    pts = np.ones((2, 50))
    a = K.lcv_bandwidths(pts)
    p0 = K.kde(pts, [1.0])
    b = np.array([K.getBW(K.ksize(K.marginal(p0, [i + 1])))[0, 0] for i in range(2)])
    assert np.array_equal(a, b)
    with pytest.raises(K.KDEError):
        K.lcv_bandwidths(np.zeros((1, 1)))


def test_no_device_memory_growth_over_many_calls():
    """200 rounds of create -> evaluate -> LOOCV -> product -> destroy leave the device's free memory where it was
    (stream-ordered pool: blocks are reused, nothing leaks)."""
    import ctypes as C
    from kde_b200 import _lib

    def free_bytes():
        fb = C.c_size_t(0)
        _lib.check(_lib.lib().kdeb200_device_props(None, None, None, None, C.byref(fb)))
        return fb.value
    rng = np.random.default_rng(5)

    def round_trip():
        p = K.kde(rng.normal(size=(2, 300)))
        q = K.kde(rng.normal(size=(2, 200)) + 1.0, [0.4, 0.5])
        K.evaluateDualTree(p, rng.normal(size=(2, 100)))
        K.entropy(q)
        K.prodAppxMSGibbsS(None, [p, q], None, None, Niter=2, Np=256, seed=1)
        p._invalidate()
        q._invalidate()
    for _ in range(20):
        round_trip()  # warm the pool
    before = free_bytes()
    for _ in range(200):
        round_trip()
    after = free_bytes()
    assert before - after < 8 << 20, (before, after)


def test_concurrent_host_threads_are_serialised():
    """The reference is single-threaded; the library promises one call at a time per process.  Four host threads
    hammering evaluation, LOOCV and Gibbs calls concurrently (ctypes drops the GIL) get the serial answers."""
    import threading
    rng = np.random.default_rng(12)
    p = K.kde(mixture(rng, 3, 500), [0.3, 0.4, 0.5])
    q = K.kde(mixture(rng, 3, 400) + 0.5, [0.4, 0.4, 0.4])
    pos = rng.normal(size=(3, 300))
    ref_e = K.evaluateDualTree(p, pos)
    ref_h = K.entropy(q)
    ref_g = K.prodAppxMSGibbsS(None, [p, q], None, None, Niter=3, Np=500, seed=3)
    ref_b = K.lcv_bandwidths(pos)
    p._dev(gibbs=True), q._dev(gibbs=True)  # handles are shared below; create them once
    errors = []

    def worker(k):
        try:
            for _ in range(15):
                assert np.array_equal(K.evaluateDualTree(p, pos), ref_e)
                assert K.entropy(q) == ref_h
                g = K.prodAppxMSGibbsS(None, [p, q], None, None, Niter=3, Np=500, seed=3)
                assert np.array_equal(g[0], ref_g[0]) and np.array_equal(g[1], ref_g[1])
                assert np.array_equal(K.lcv_bandwidths(pos), ref_b)
                with pytest.raises(K.KDEError):  # the error channel is per thread
                    K.evaluateDualTree(p, np.zeros((2, 3)))
        except Exception as e:  # noqa: BLE001
            errors.append((k, repr(e)))
    th = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors


def test_sharded_lcv_entry_point_single_process_and_callback_errors():
    """kdeb200_kde_lcv_sharded with one process owning all rows == kdeb200_kde_lcv (large-N host loop); a failing
    all-reduce callback surfaces as an error instead of unwinding through the C frames."""
    import ctypes as C
    from kde_b200 import _lib, dist as kd
    rng = np.random.default_rng(77)
    pts = mixture(rng, 2, 1500)
    assert np.array_equal(K.getBW(kd.kde_sharded(pts))[:, 0], K.getBW(K.kde(pts))[:, 0])
    calls = []

    def two_halves(ps, pf, _user):  # pretend a second process contributed nothing
        calls.append(ps[0])
        return 0
    flat = np.ascontiguousarray(pts.T).ravel()
    bw = np.zeros(2)
    cb = _lib.allreduce_fn(two_halves)
    _lib.check(_lib.lib().kdeb200_kde_lcv_sharded(2, 1500, _lib.fptr(flat), 0, 1500, cb, None, _lib.fptr(bw), None))
    assert len(calls) >= 20 and np.array_equal(bw, K.lcv_bandwidths(pts))
    bad = _lib.allreduce_fn(lambda ps, pf, u: 5)
    rc = _lib.lib().kdeb200_kde_lcv_sharded(2, 1500, _lib.fptr(flat), 0, 1500, bad, None, _lib.fptr(bw), None)
    assert rc == 9 and b"all-reduce callback failed" in _lib.lib().kdeb200_last_error()
    rc = _lib.lib().kdeb200_kde_lcv_sharded(2, 1500, _lib.fptr(flat), 10, 5, cb, None, _lib.fptr(bw), None)
    assert rc == 3


def test_update_bandwidth_invalidates_the_device_records():
    """updateBandwidth! (src/CrossValidation.jl:5-12) followed by evaluation / LOO / a product uses the NEW
    bandwidth, like the reference: the cached device handle is rebuilt."""
    rng = np.random.default_rng(41)
    pts = mixture(rng, 2, 300)
    p = K.kde(pts, [0.3, 0.4])
    pos = rng.normal(size=(2, 64))
    before = K.evaluateDualTree(p, pos)
    K.updateBandwidth(p, p.bandwidth * 2.25)              # std x 1.5 on every node
    fresh = K.kde(pts, [0.45, 0.6])
    after = K.evaluateDualTree(p, pos)
    assert relerr(after, K.evaluateDualTree(fresh, pos)) < 1e-13 and relerr(after, before) > 1e-3
    assert relerr(K.evaluateDualTree(p, p), K.evaluateDualTree(fresh, fresh)) < 1e-13
    assert relerr(p(pos), K.evaluateDualTree(fresh, pos)) < 1e-13
    assert abs(K.entropy(p) - K.entropy(fresh)) < 1e-13 * abs(K.entropy(fresh))
    # Gibbs reads the INTERNAL nodes' variances too, and updateBandwidth! scales those as well (2.25 x (h^2 + spread),
    # not what kde!(pts, 1.5 h) builds): compare with a density object assembled from the very same scaled arrays
    same = K.BallTreeDensity(p.bt, p.means.copy(), p.bandwidth.copy())
    g1 = K.prodAppxMSGibbsS(None, [p, p], None, None, Niter=2, Np=128, seed=9)
    g2 = K.prodAppxMSGibbsS(None, [same, same], None, None, Niter=2, Np=128, seed=9)
    g3 = K.prodAppxMSGibbsS(None, [fresh, fresh], None, None, Niter=2, Np=128, seed=9)
    assert np.array_equal(g1[1], g2[1]) and np.array_equal(g1[0], g2[0]) and not np.array_equal(g1[1], g3[1])
    # nLOO_LL's multiply / divide-back leaves the (ulp-drifted) bandwidth in place and later calls see it
    H = K.nLOO_LL(1.3, p)
    o = OKDE.kde_bw(pts, [0.45, 0.6])
    assert abs(H - o.nloo_ll(1.3)) <= 1e-12 * abs(H)
    assert relerr(K.evaluateDualTree(p, pos), o.evaluate(pos)) < TOL   # the oracle drifted the same way
    with pytest.raises(K.KDEError):
        K.updateBandwidth(p, p.bandwidth[:10])


def test_fp32_refuses_data_too_wide_for_its_contract():
    """FP32 coordinates are centred and scaled to bandwidth units; beyond extent/sigma ~ 2000 their rounding breaks
    the 1e-5 contract, so the library refuses (no silent accuracy loss, no hidden FP64 fallback)."""
    rng = np.random.default_rng(2)
    pts = np.hstack([rng.normal(size=(1, 200)), 5000.0 + rng.normal(size=(1, 200))])
    p = K.kde(pts, [0.5])
    x = pts[:, ::7]
    assert relerr(K.evaluateDualTree(p, x), OKDE.kde_bw(pts, [0.5]).evaluate(x)) < TOL
    with pytest.raises(K.KDEError, match="FP32"):
        K.evaluateDualTree(p, x, precision=K.F32)
    q = K.kde(pts, [5.0])                                  # the same data at a 10x wider bandwidth is fine
    assert relerr(K.evaluateDualTree(q, x, precision=K.F32), K.evaluateDualTree(q, x)) < 1e-5


def test_full_size_c3_loo_rows_and_kde():
    """BASELINE config 3 at full size (kde! LOOCV of 100 000 x 4-D mixture points).
    (a) One marginal's LOO sum at 100k rows -- component splits S > 1, eval_finalize_kernel and the 100k-row
        likelihood reduction, which no smaller test reaches -- against the oracle's literal evalDirect rows
        (okde_loo_rows) on 512 rows spread over the leaf range: 1e-12, through the full LOO evaluation and
        through kdeb200_loo_partial(j0, j1).
    (b) The whole kde!(points): the native loop equals the step-by-step mirror bit for bit on one dimension, and the
        selected bandwidths are where the normal-reference rule puts them for this mixture."""
    import ctypes as C
    from kde_b200 import _lib
    rng = np.random.default_rng(20261017)
    N = 100_000
    pts = mixture(rng, 4, N)
    # (a) marginal of dimension 1 at a mid-bracket bandwidth
    x = pts[:1]
    h = 0.05
    p, o = K.kde(x, [h]), OKDE.kde_bw(x, [h])
    L = K.evaluateDualTree(p, p)                            # original order
    perm = p.bt.permutation[N:] - 1                         # leaf -> original index
    blocks = [(0, 128), (33_333, 33_333 + 128), (77_777, 77_777 + 128), (N - 128, N)]
    s_tot = 0.0
    for a, b in blocks:
        exp = o.loo_rows(a, b, nthreads=8)
        assert relerr(L[perm[a:b]], exp) < TOL
        s, f = C.c_double(0), C.c_int(0)
        bw = np.array([h * h])
        _lib.check(_lib.lib().kdeb200_loo_partial(p._dev(), _lib.fptr(bw), a, b, C.byref(s), C.byref(f)))
        w = p.bt.weights[N + a:N + b]
        ref = float(np.sum(np.log(exp) * w))
        assert f.value == 0 and abs(s.value - ref) <= 1e-12 * abs(ref)
        s_tot += s.value
    H = K.entropy(p)
    w_all = p.bt.weights[N:]
    assert abs(H + float(np.sum(np.log(L[perm]) * w_all))) <= 1e-11 * abs(H)   # reduction == sum of its rows
    # (b) the full kde!(points): native loop == step-by-step mirror (dimension 1), and sane, positive bandwidths
    calls = []
    bw = K.lcv_bandwidths(pts, _count=calls)
    assert bw.shape == (4,) and np.all(bw > 0.005) and np.all(bw < 0.5) and all(10 <= c <= 40 for c in calls)
    cnt = []
    pp = K.ksize(K.marginal(K.kde(pts, [1.0]), [1]), _count=cnt)
    assert K.getBW(pp)[0, 0] == bw[0] and cnt[0] == calls[0]
    # and the selected bandwidths sit where theory puts them for this mixture (two modes of sigma 0.6 per dimension,
    # 50k points each: 1.06 sigma n^(-1/5) = 0.073)
    assert np.all(np.abs(bw - 0.073) < 0.012)


def test_known_constructions_through_the_gpu_path():
    """test/testKnownConstructions.jl "should get" comments (tree arrays of test01/02, the whole kde!(3 points) LOOCV
    result of test03: leaf variance 0.038521) through the library: host builder + device LOOCV."""
    from tests.test_oracle_golden import check_known_constructions

    def arrays(p):
        return {"centers": p.bt.centers, "ranges": p.bt.ranges, "weights": p.bt.weights, "means": p.means,
                "bandwidth": p.bandwidth, "highest_leaf": p.bt.highest_leaf, "lowest_leaf": p.bt.lowest_leaf,
                "permutation": p.bt.permutation}
    check_known_constructions(lambda pts, bw, w: K.kde(pts, bw, w), lambda pts: K.kde(pts), arrays)
