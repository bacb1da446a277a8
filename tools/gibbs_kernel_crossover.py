"""Thread-per-chain vs warp-per-chain Gibbs kernel by number of samples (kernel ms from the library's event bracket).
usage: python tools/gibbs_kernel_crossover.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kde_b200 as K

K.init(0)
rng = np.random.default_rng(1)
shapes = {
    "c1 (2 x 100, 2-D)": [K.kde(rng.standard_normal((2, 100)) + 2 * j) for j in range(2)],
    "c2 (300 + 100, 1-D)": [K.kde(rng.beta(1.0, 0.45, size=(1, 300))), K.kde(rng.rayleigh(0.5, size=(1, 100)) - 0.5)],
    "6 x 100, 3-D": [K.kde(rng.standard_normal((3, 100))) for j in range(6)],
    "4 x 1000, 3-D": [K.kde(rng.standard_normal((3, 1000)), [0.3]) for j in range(4)],
}
out = {}
for name, trees in shapes.items():
    rows = []
    for Np in (100, 1000, 4000, 10000, 20000, 40000, 80000):
        r = {"Np": Np}
        for label, lim in (("thread_ms", "0"), ("warp_ms", "1000000000")):
            os.environ["KDEB200_GIBBS_WARP_MAX"] = lim
            K.prodAppxMSGibbsS(None, trees, None, None, Niter=5, Np=Np, seed=1)
            ms = []
            for rep in range(3):
                K.prodAppxMSGibbsS(None, trees, None, None, Niter=5, Np=Np, seed=rep)
                ms.append(K.last_kernel_ms()[0])
            r[label] = float(np.median(ms))
        rows.append(r)
    out[name] = rows
print(json.dumps(out, indent=1))
