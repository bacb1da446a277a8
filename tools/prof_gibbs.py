"""Runs the C4-shaped Gibbs kernel once or twice (for ncu): python tools/prof_gibbs.py [samples] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kde_b200 as K
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 75776
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
K.init(0)
trees = [K.kde(bench.synth_points(j), bench.silverman(bench.synth_points(j))) for j in range(bench.NDENS)]
for r in range(reps):
    p, i = K.prodAppxMSGibbsS(None, trees, None, None, Niter=bench.NITER, Np=n, seed=bench.SEED)
    ms, nl = K.last_kernel_ms()
    print("rep", r, "samples", n, "kernel ms", ms, "samples/s", n / ms * 1e3)
