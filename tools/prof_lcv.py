"""Runs the fused golden-section LOOCV kernel once or twice (for ncu): python tools/prof_lcv.py [N] [d]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kde_b200 as K
N = int(sys.argv[1]) if len(sys.argv) > 1 else 400
d = int(sys.argv[2]) if len(sys.argv) > 2 else 3
K.init(0)
pts = np.random.default_rng(3).standard_normal((d, N))
for r in range(2):
    calls = []
    bw = K.lcv_bandwidths(pts, _count=calls)
    print("bw", bw, "nLOO_LL calls", calls, "kernel ms", K.last_kernel_ms())
