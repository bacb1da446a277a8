"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/kdeb200.h declares (no compute calls: there is no GPU here), the ctypes table matches
the header, and host-only entry points (tree build) agree with the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "kdeb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(kdeb200_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g._load_build().build()
    from kde_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol(built):
    L = ctypes.CDLL(built.SO_PATH)
    names = header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), n


def test_ctypes_table_matches_header(built):
    assert sorted(built.SIGNATURES) == header_symbols()


def test_error_channel_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import kde_b200 as K
    p = K.kde(np.arange(6.0).reshape(2, 3), [1.0])  # host-only: fine
    with pytest.raises(K.KDEError):  # no CPU fallback: the device path fails loudly
        K.evaluateDualTree(p, np.zeros((2, 4)))


def test_gibbs_precision_switch_is_validated(built):
    """kdeb200_set_gibbs_precision is a process-wide host setting (no device needed): F64 / F32 accepted, the evaluation-only
    modes rejected; the switch never reaches a CPU path -- without a GPU the sampler still fails loudly."""
    import kde_b200 as K
    for bad in (K.F64_BOUNDED, K.F32_BOUNDED, 7, -1):
        with pytest.raises(K.KDEError):
            K.set_gibbs_precision(bad)
    K.set_gibbs_precision(K.F32)
    K.set_gibbs_precision(K.F64)
    import torch
    if not torch.cuda.is_available():
        p = K.kde(np.arange(12.0).reshape(2, 6), [1.0])
        K.set_gibbs_precision(K.F32)
        try:
            with pytest.raises(K.KDEError):
                K.prodAppxMSGibbsS(None, [p, p], None, None, Niter=1, Np=5000, seed=1)
        finally:
            K.set_gibbs_precision(K.F64)


@pytest.mark.parametrize("d,N", [(1, 1), (1, 2), (1, 4), (2, 3), (3, 100), (4, 257), (2, 1024), (8, 33),
                                 (3, 70001), (1, 140000)])  # the last two take the multi-threaded path
def test_host_tree_builder_equals_oracle(built, d, N):
    import kde_b200 as K
    from oracle.oracle import OKDE
    rng = np.random.default_rng(d * 1000 + N)
    pts = rng.standard_normal((d, N))
    pts[:, N // 2:] = np.round(pts[:, N // 2:], 1)  # ties
    bw, w = rng.random(d) + 0.1, rng.random(N) + 0.1
    a, o = K.kde(pts, bw, w), OKDE.kde_bw(pts, bw, w).arrays()
    for k in ("centers", "ranges", "weights", "left_child", "right_child", "lowest_leaf", "highest_leaf", "permutation"):
        assert np.array_equal(getattr(a.bt, k), o[k]), k
    for k in ("means", "bandwidth", "bandwidthMin", "bandwidthMax"):
        assert np.array_equal(getattr(a, k), o[k]), k
    assert np.array_equal(K.getPoints(a), pts)
    assert np.allclose(K.getBW(a)[:, 0], bw, rtol=1e-15) and np.allclose(K.getWeights(a), w / w.sum(), rtol=1e-14)


def test_host_api_mirrors(built):
    import kde_b200 as K
    from oracle.oracle import OKDE
    rng = np.random.default_rng(5)
    pts = rng.standard_normal((3, 50))
    p, o = K.kde(pts, [0.2, 0.3, 0.4]), OKDE.kde_bw(pts, [0.2, 0.3, 0.4])
    m, om = K.marginal(p, [1, 3]), o.marginal([1, 3])
    assert np.array_equal(m.means, om.arrays()["means"]) and np.array_equal(m.bandwidth, om.arrays()["bandwidth"])
    assert K.neighborMinMax(p) == pytest.approx(o.neighbor_minmax(), rel=1e-15)
    import torch
    if torch.cuda.is_available():
        s, idx = K.sample(p, 20, rng=np.random.default_rng(1))
        assert s.shape == (3, 20) and idx.min() >= 1 and idx.max() <= 50
    else:  # sample draws on the device (kdeb200_sample): without one it fails loudly, there is no CPU fallback
        with pytest.raises(K.KDEError):
            K.sample(p, 20, rng=np.random.default_rng(1))
    with pytest.raises(K.KDEError):
        K.kde(pts, [0.1, 0.2])
    with pytest.raises(K.KDEError):
        K.kde(pts, [0.1], addop=(lambda a, b: a + b,))


def test_header_is_plain_c(tmp_path):
    """include/kdeb200.h must be consumable by a C compiler (the Julia ccall / cgo / JNI side sees C)."""
    import subprocess
    src = tmp_path / "use.c"
    src.write_text('#include "kdeb200.h"\nint main(void) { kdeb200_tree_t t = 0; int n = 0; (void)t;\n'
                   '  return kdeb200_device_count(&n) == 0 ? kdeb200_version() - kdeb200_version() : 0; }\n')
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           "-c", str(src), "-o", str(tmp_path / "use.o")])


def test_host_wrappers_without_device(built):
    """String round trip (test/runtests.jl:246-255), ranges, mean, fit: pure host code."""
    import kde_b200 as K
    rng = np.random.default_rng(2)
    p = K.kde(rng.standard_normal((2, 3)), [0.3, 0.4])
    pp = K.from_string(K.to_string(p))
    assert np.linalg.norm(K.getPoints(pp) - K.getPoints(p)) < 1e-4 and np.linalg.norm(K.getBW(pp) - K.getBW(p)) < 1e-4
    r = K.getKDERange(p, extend=0.1)
    pts = K.getPoints(p)
    assert np.allclose(r[:, 0], pts.min(1) - 0.1 * (pts.max(1) - pts.min(1)))
    assert np.allclose(K.getKDEMean(p), pts.mean(1))
    mu, c = K.getKDEfit(p)
    assert np.allclose(mu, pts.mean(1)) and c.shape == (2, 2)
    with pytest.raises(K.KDEError):
        K.from_string("nope")


def _build_c_example(tmp_path, built):
    import subprocess
    exe = tmp_path / "product_c"
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "product_c.c"), "-o", str(exe),
                           "-L", os.path.dirname(built.SO_PATH), "-lkdeb200", "-lm",
                           "-Wl,-rpath," + os.path.dirname(built.SO_PATH)])
    return exe


def test_c_example_links_against_the_abi(built, tmp_path):
    """examples/product_c.c: plain C99 against include/kdeb200.h + libkdeb200.so (no Python / torch)."""
    import subprocess
    import torch
    exe = _build_c_example(tmp_path, built)
    if not torch.cuda.is_available():
        r = subprocess.run([str(exe)], capture_output=True, text=True)
        assert r.returncode in (1, 2) and "no CUDA" in (r.stderr + r.stdout) or "cuda" in (r.stderr + r.stdout).lower()


@pytest.mark.gpu
def test_c_example_runs_on_gpu(built, tmp_path):
    import subprocess
    exe = _build_c_example(tmp_path, built)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "product mean" in r.stdout
