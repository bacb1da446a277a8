"""FP32 evaluation kernel time at C5-like sizes: python tools/bench_f32.py [N]   (KDEB200_SO selects a variant)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kde_b200 as K
from tests.util import mixture, silverman
N = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
K.init(0)
rng = np.random.default_rng(1)
for d in (1, 2, 3, 4, 6):
    pts, pos = mixture(rng, d, N), mixture(rng, d, N)
    p = K.kde(pts, silverman(pts))
    ref = K.evaluateDualTree(p, pos[:, :2000])
    K.evaluateDualTree(p, pos, precision=K.F32)
    got = K.evaluateDualTree(p, pos, precision=K.F32)
    ms, _ = K.last_kernel_ms()
    err = float(np.max(np.abs(got[:2000] - ref) / ref))
    print("d=%d f32 eval %.2f ms (%.3g evals/s) max rel err %.2e" % (d, ms, N * N / ms * 1e3, err))
