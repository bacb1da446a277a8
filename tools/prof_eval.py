"""Runs the FP64 / FP32 evaluation and the LOO likelihood once each (for ncu): python tools/prof_eval.py [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kde_b200 as K
from tests.util import mixture, silverman
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
K.init(0)
rng = np.random.default_rng(1)
pts, pos = mixture(rng, 3, N), mixture(rng, 3, N)
p = K.kde(pts, silverman(pts))
for prec in (K.F64, K.F32):
    K.evaluateDualTree(p, pos, precision=prec)
    print("eval", prec, K.last_kernel_ms())
p1 = K.marginal(p, [1])
print("entropy", K.entropy(p1), K.last_kernel_ms())
