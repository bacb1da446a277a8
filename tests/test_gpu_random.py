"""-m gpu: randomised shapes (hypothesis) for the three entry points, against the oracle."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import kde_b200 as K
from oracle import oracle as O
from tests.util import relerr

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
SET = dict(max_examples=25, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)


@settings(**SET)
@given(d=st.integers(1, 8), N=st.integers(1, 400), M=st.integers(1, 300), seed=st.integers(0, 2 ** 31),
       scale=st.sampled_from([1e-3, 1.0, 1e4]), offset=st.sampled_from([0.0, 1e3]))
def test_eval_random_shapes(d, N, M, seed, scale, offset):
    rng = np.random.default_rng(seed)
    pts = offset + scale * rng.standard_normal((d, N))
    bw = scale * (0.2 + rng.random(d))
    w = rng.random(N) + 1e-3
    pos = offset + scale * 1.5 * rng.standard_normal((d, M))
    p, o = K.kde(pts, bw, w), O.OKDE.kde_bw(pts, bw, w)
    exp = o.evaluate(pos)
    got = K.evaluateDualTree(p, pos)
    nz = exp > 1e-290
    assert np.all(got[~nz] <= 1e-289)
    # 1e-12 is the contract for well-conditioned data; data sitting 1e3 bandwidths from the origin loses
    # digits in x - mu itself (both sides), so the bound scales with |x|/bandwidth
    tol = 1e-12 * max(1.0, offset / scale / 10.0)
    assert relerr(got[nz], exp[nz]) < tol
    if N > 1:
        assert relerr(K.evaluateDualTree(p, p), o.evaluate()) < max(tol, 1e-12)


@settings(**SET)
@given(d=st.integers(1, 8), M=st.integers(1, 6), N=st.integers(1, 150), Np=st.integers(1, 90), T=st.integers(0, 6),
       seed=st.integers(0, 2 ** 31), ent=st.booleans())
def test_gibbs_random_shapes(d, M, N, Np, T, seed, ent):
    rng = np.random.default_rng(seed)
    Ns = [max(1, int(N * f)) for f in rng.uniform(0.3, 1.0, size=M)]
    pts = [rng.standard_normal((d, n)) + 0.3 * j for j, n in enumerate(Ns)]
    bws = [0.2 + rng.random(d) for _ in range(M)]
    ws = [rng.random(n) + 0.01 for n in Ns]
    kt = [K.kde(p, b, w) for p, b, w in zip(pts, bws, ws)]
    ot = [O.OKDE.kde_bw(p, b, w) for p, b, w in zip(pts, bws, ws)]
    nU, nN = O.prod_sizes(ot, Np, T)
    U, G = rng.random(nU), rng.standard_normal(nN)
    ep, ei = O.gibbs(ot, Np, T, U, G, add_entropy=ent)
    gp, gi = K.prodAppxMSGibbsS(None, kt, None, None, Niter=T, Np=Np, randU=U, randN=G, addEntropy=ent)
    assert np.array_equal(gi, ei)
    assert np.max(np.abs(gp - ep)) <= 1e-10 * max(1.0, np.max(np.abs(ep)))
