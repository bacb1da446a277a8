"""Kernel time of the FP64 evaluation / LOO kernels per dimension (tuning of queries per thread):
python tools/bench_eval_dims.py [N]   (KDEB200_SO selects a variant build)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kde_b200 as K
from tests.util import mixture, silverman
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
K.init(0)
rng = np.random.default_rng(1)
for d in range(1, 9):
    pts, pos = mixture(rng, d, N), mixture(rng, d, N)
    p = K.kde(pts, silverman(pts))
    K.evaluateDualTree(p, pos)
    K.evaluateDualTree(p, pos)
    ms, _ = K.last_kernel_ms()
    K.entropy(p)
    K.entropy(p)
    ms2, _ = K.last_kernel_ms()
    print("d=%d eval %.2f ms (%.3g evals/s)  loo %.2f ms" % (d, ms, N * N / ms * 1e3, ms2))
