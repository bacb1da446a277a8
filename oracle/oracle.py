"""ctypes binding of the CPU oracle (oracle/kde_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

f64p = C.POINTER(C.c_double)
i64p = C.POINTER(C.c_int64)
u8p = C.POINTER(C.c_uint8)


def build(force=False):
    """Compile oracle/liboracle.so with the committed Makefile (building the checker is not using it)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("kde_oracle.c", "kde_oracle.h", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.okde_make.restype = C.c_void_p
        L.okde_make.argtypes = [C.c_int64, C.c_int64, f64p, f64p, f64p]
        L.okde_kde_bw.restype = C.c_void_p
        L.okde_kde_bw.argtypes = [C.c_int64, C.c_int64, f64p, f64p, C.c_int64, f64p]
        L.okde_kde_lcv.restype = C.c_void_p
        L.okde_kde_lcv.argtypes = [C.c_int64, C.c_int64, f64p, i64p]
        L.okde_free.argtypes = [C.c_void_p]
        for n in ("okde_get_points", "okde_get_bw", "okde_get_weights"):
            getattr(L, n).argtypes = [C.c_void_p, f64p]
        L.okde_marginal.restype = C.c_void_p
        L.okde_marginal.argtypes = [C.c_void_p, i64p, C.c_int64]
        L.okde_dims.restype = C.c_int64
        L.okde_dims.argtypes = [C.c_void_p]
        L.okde_npts.restype = C.c_int64
        L.okde_npts.argtypes = [C.c_void_p]
        L.okde_arr_f.restype = f64p
        L.okde_arr_f.argtypes = [C.c_void_p, C.c_int]
        L.okde_arr_i.restype = i64p
        L.okde_arr_i.argtypes = [C.c_void_p, C.c_int]
        L.okde_evaluate.argtypes = [C.c_void_p, C.c_void_p, f64p]
        L.okde_eval_points.argtypes = [C.c_void_p, C.c_int64, f64p, f64p]
        L.okde_eval_points_omp.argtypes = [C.c_void_p, C.c_int64, f64p, f64p, C.c_int]
        L.okde_eval_avg_logl.restype = C.c_double
        L.okde_eval_avg_logl.argtypes = [C.c_void_p, C.c_void_p]
        L.okde_entropy.restype = C.c_double
        L.okde_entropy.argtypes = [C.c_void_p]
        L.okde_nloo_ll.restype = C.c_double
        L.okde_nloo_ll.argtypes = [C.c_double, C.c_void_p]
        L.okde_neighbor_minmax.argtypes = [C.c_void_p, f64p, f64p]
        L.okde_ksize.restype = C.c_void_p
        L.okde_ksize.argtypes = [C.c_void_p, i64p]
        L.okde_gibbs_nlevels.restype = C.c_int64
        L.okde_gibbs_nlevels.argtypes = [C.POINTER(C.c_void_p), C.c_int64]
        L.okde_gibbs.argtypes = [C.c_int64, C.POINTER(C.c_void_p), C.c_int64, C.c_int64, f64p, i64p, f64p,
                                 C.c_int64, f64p, C.c_int64, C.c_int, u8p, C.c_int64, C.c_int64]
        L.okde_gibbs_record.argtypes = L.okde_gibbs.argtypes + [i64p]
        L.okde_gibbs_omp.argtypes = [C.c_int64, C.POINTER(C.c_void_p), C.c_int64, C.c_int64, f64p, i64p, f64p,
                                     C.c_int64, f64p, C.c_int64, C.c_int, u8p, C.c_int]
        L.okde_max_threads.restype = C.c_int
        L.okde_loo_rows.argtypes = [C.c_void_p, C.c_int64, C.c_int64, f64p, C.c_int]
        L.okde_set_sum_simd.argtypes = [C.c_int, C.c_int]
        _LIB = L
    return _LIB


def _f(a):
    return a.ctypes.data_as(f64p)


def _i(a):
    return a.ctypes.data_as(i64p)


def _colmajor(points):
    """Reference matrices are d x N column-major; numpy callers pass shape (d, N)."""
    p = np.asarray(points, dtype=np.float64)
    if p.ndim == 1:
        p = p.reshape(1, -1)
    return np.asfortranarray(p)


class OKDE:
    """Owning handle on an oracle BallTreeDensity."""

    def __init__(self, ptr):
        if not ptr:
            raise RuntimeError("oracle returned NULL")
        self.ptr = C.c_void_p(ptr)

    def __del__(self):
        try:
            if self.ptr:
                lib().okde_free(self.ptr)
                self.ptr = None
        except Exception:
            pass

    # -- constructors mirroring kde!(...) ------------------------------------------
    @staticmethod
    def kde_bw(points, ks, weights=None):
        p = _colmajor(points)
        d, N = p.shape
        ks = np.atleast_1d(np.asarray(ks, dtype=np.float64)).ravel()
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        return OKDE(lib().okde_kde_bw(d, N, _f(p), _f(ks), len(ks), None if w is None else _f(w)))

    @staticmethod
    def kde_lcv(points):
        p = _colmajor(points)
        d, N = p.shape
        n = C.c_int64(0)
        o = OKDE(lib().okde_kde_lcv(d, N, _f(p), C.byref(n)))
        o.n_loo_calls = n.value
        return o

    # -- accessors ---------------------------------------------------------------------
    @property
    def dims(self):
        return lib().okde_dims(self.ptr)

    @property
    def npts(self):
        return lib().okde_npts(self.ptr)

    def arrays(self):
        """Copies of every tree array, named as in the reference structs."""
        d, N = self.dims, self.npts
        L = lib()
        fl = {"centers": (0, 2 * N * d), "ranges": (1, 2 * N * d), "weights": (2, 2 * N), "means": (3, 2 * N * d),
              "bandwidth": (4, 2 * N * d), "bandwidthMin": (5, N * d), "bandwidthMax": (6, N * d)}
        il = {"left_child": 0, "right_child": 1, "lowest_leaf": 2, "highest_leaf": 3, "permutation": 4}
        out = {"dims": d, "num_points": N}
        for k, (w, n) in fl.items():
            out[k] = np.ctypeslib.as_array(L.okde_arr_f(self.ptr, w), shape=(n,)).copy()
        for k, w in il.items():
            out[k] = np.ctypeslib.as_array(L.okde_arr_i(self.ptr, w), shape=(2 * N,)).copy()
        return out

    def get_points(self):
        out = np.zeros((self.dims, self.npts), order="F")
        lib().okde_get_points(self.ptr, _f(out))
        return out

    def get_bw(self):
        out = np.zeros((self.dims, self.npts), order="F")
        lib().okde_get_bw(self.ptr, _f(out))
        return out

    def get_weights(self):
        out = np.zeros(self.npts)
        lib().okde_get_weights(self.ptr, _f(out))
        return out

    def marginal(self, ind):
        ind = np.ascontiguousarray(ind, dtype=np.int64)
        return OKDE(lib().okde_marginal(self.ptr, _i(ind), len(ind)))

    # -- evaluation --------------------------------------------------------------------
    def evaluate(self, pos=None, nthreads=0):
        """pos=None => leave-one-out self evaluation (evaluate(bd, bd, ...))."""
        if pos is None:
            out = np.zeros(self.npts)
            lib().okde_evaluate(self.ptr, self.ptr, _f(out))
            return out
        if isinstance(pos, OKDE):
            out = np.zeros(pos.npts)
            rc = lib().okde_evaluate(self.ptr, pos.ptr, _f(out))
            if rc:
                raise RuntimeError("evaluate -- dimensions of two BallTreeDensities must match")
            return out
        p = _colmajor(pos)
        if p.shape[0] != self.dims:
            raise RuntimeError("bd and pos must have the same dimension")
        out = np.zeros(p.shape[1])
        if nthreads:
            lib().okde_eval_points_omp(self.ptr, p.shape[1], _f(p), _f(out), nthreads)
        else:
            lib().okde_eval_points(self.ptr, p.shape[1], _f(p), _f(out))
        return out

    def loo_rows(self, j0, j1, nthreads=0):
        """LOO densities of the leaf rows [j0, j1) in LEAF order (evaluate(bd, bd) restricted to a block)."""
        out = np.zeros(j1 - j0)
        if lib().okde_loo_rows(self.ptr, j0, j1, _f(out), int(nthreads)):
            raise RuntimeError("loo_rows: bad row range")
        return out

    def eval_avg_logl(self, other):
        return lib().okde_eval_avg_logl(self.ptr, other.ptr)

    def entropy(self):
        return lib().okde_entropy(self.ptr)

    def nloo_ll(self, alpha):
        return lib().okde_nloo_ll(float(alpha), self.ptr)

    def neighbor_minmax(self):
        a, b = C.c_double(0), C.c_double(0)
        lib().okde_neighbor_minmax(self.ptr, C.byref(a), C.byref(b))
        return a.value, b.value

    def ksize(self):
        n = C.c_int64(0)
        o = OKDE(lib().okde_ksize(self.ptr, C.byref(n)))
        o.n_loo_calls = n.value
        return o


def gibbs_nlevels(trees):
    arr = (C.c_void_p * len(trees))(*[t.ptr for t in trees])
    return lib().okde_gibbs_nlevels(arr, len(trees))


def gibbs(trees, Np, Niter, randU, randN, add_entropy=True, mask=None, s0=0, s1=None, nthreads=0, record=False):
    """gibbs1 with injected random streams.  Returns (points d x Np, indices M x Np) and, with
    record=True, labelsChoosen as an int64 array [Np, M, Nlevels] (-1 = never written)."""
    M = len(trees)
    d = max(t.dims for t in trees)
    arr = (C.c_void_p * M)(*[t.ptr for t in trees])
    pts = np.zeros((d, Np), order="F")
    ind = np.ones((M, Np), dtype=np.int64, order="F")
    randU = np.ascontiguousarray(randU, dtype=np.float64)
    randN = np.ascontiguousarray(randN, dtype=np.float64)
    mk = None
    if mask is not None:
        mk = np.ascontiguousarray(np.asarray(mask, dtype=np.uint8).reshape(M, d))
    mp = None if mk is None else mk.ctypes.data_as(u8p)
    if record:
        rec = np.full((Np, M, gibbs_nlevels(trees)), -1, dtype=np.int64)
        rc = lib().okde_gibbs_record(M, arr, Np, Niter, _f(pts), _i(ind), _f(randU), randU.size, _f(randN), randN.size,
                                     int(bool(add_entropy)), mp, s0, Np if s1 is None else s1, _i(rec))
        if rc:
            raise RuntimeError("oracle gibbs failed rc=%d" % rc)
        return pts, ind, rec
    if nthreads:
        rc = lib().okde_gibbs_omp(M, arr, Np, Niter, _f(pts), _i(ind), _f(randU), randU.size, _f(randN), randN.size,
                                  int(bool(add_entropy)), mp, int(nthreads))
    else:
        rc = lib().okde_gibbs(M, arr, Np, Niter, _f(pts), _i(ind), _f(randU), randU.size, _f(randN), randN.size,
                              int(bool(add_entropy)), mp, s0, Np if s1 is None else s1)
    if rc:
        raise RuntimeError("oracle gibbs failed rc=%d" % rc)
    return pts, ind


def prod_sizes(trees, Np, Niter):
    """Stream sizes allocated by prodAppxMSGibbsS (src/MSGibbs01.jl:656-662)."""
    import math
    M = len(trees)
    d = max(t.dims for t in trees)
    maxNp = max([Np] + [t.npts for t in trees])
    nlev = int(math.floor(math.log(float(maxNp)) / math.log(2.0) + 1.0))
    return Np * M * (Niter + 2) * nlev, d * Np * (nlev + 1)


def set_sum_simd(vf=0, ic=0):
    """Emulated shape of Julia's @simd sum(lambdas) for Ndens >= 16 (0, 0 = sequential)."""
    lib().okde_set_sum_simd(int(vf), int(ic))


def max_threads():
    return lib().okde_max_threads()
