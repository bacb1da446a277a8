"""Per-call latency of the small-N paths (robotics-sized KDEs): entropy, evaluate, kde!, product."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kde_b200 as K
K.init(0)
rng = np.random.default_rng(0)
def med(f, n=50):
    f(); ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return 1e6 * float(np.median(ts))
pts = rng.standard_normal((2, 100))
p = K.kde(pts, [0.3]); p._dev()
x = rng.standard_normal((2, 100))
print("entropy(N=100)        %.0f us" % med(lambda: K.entropy(p)))
print("evaluate(100x100)     %.0f us" % med(lambda: K.evaluateDualTree(p, x)))
print("kde(pts,bw) host build %.0f us" % med(lambda: K.kde(pts, [0.3])))
def cd():
    q = K.kde(pts, [0.3]); q._dev(); q._invalidate()
print("build+create+destroy  %.0f us" % med(cd))
print("kde!(pts) LOOCV       %.0f us" % med(lambda: K.kde(pts), 10))
q = K.kde(2 + rng.standard_normal((2, 100)), [0.3])
print("prodAppx(100 samples) %.0f us" % med(lambda: K.prodAppxMSGibbsS(None, [p, q], None, None, Niter=5, Np=100, seed=1)))
print("p*q                   %.0f us" % med(lambda: p * q, 10))
