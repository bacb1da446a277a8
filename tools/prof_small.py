"""The small-call kernels once each (for ncu): warp-per-chain Gibbs at C1, the fused product (Gibbs + on-chip LOOCV)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kde_b200 as K
K.init(0)
rng = np.random.default_rng(1)
trees = [K.kde(rng.standard_normal((2, 100)) + 2 * j) for j in range(2)]
for r in range(2):
    K.prodAppxMSGibbsS(None, trees, None, None, Niter=5, Np=100, seed=r)
    print("gibbs C1", K.last_kernel_ms())
    pq = K.prod(trees, seed=r)
    print("p*q", K.last_kernel_ms(), K.getBW(pq)[:, 0])
