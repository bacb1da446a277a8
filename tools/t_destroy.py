import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np
import kde_b200 as K
K.init(0)
rng = np.random.default_rng(0)
for N in (1000, 100000, 100000, 100000):
    p = K.kde(rng.standard_normal((1, N)), [0.1])
    t0 = time.perf_counter(); p._dev(); t1 = time.perf_counter()
    H = K.entropy(p); t2 = time.perf_counter()
    p._invalidate(); t3 = time.perf_counter()
    print(N, "create %.2f ms  entropy %.2f ms  destroy %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
