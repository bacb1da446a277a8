"""Gibbs kernel throughput per dimension at the C4 shape (8 x 4096 components, Niter=5): python tools/bench_gibbs_dims.py [Np]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kde_b200 as K
from tests.util import mixture, silverman
Np = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
K.init(0)
rng = np.random.default_rng(2)
for d in range(1, 9):
    trees = []
    for j in range(8):
        p = mixture(rng, d, 4096, 0.25 * j)
        trees.append(K.kde(p, silverman(p)))
    for r in range(2):
        K.prodAppxMSGibbsS(None, trees, None, None, Niter=5, Np=Np, seed=1)
    ms, _ = K.last_kernel_ms()
    L, pu, pn, ev = K.gibbs_sizes(trees, 5)
    print("d=%d  %.1f ms  %.3g samples/s  %.3g evals/s" % (d, ms, Np / ms * 1e3, Np * ev / ms * 1e3))
