"""Runs each evaluation kernel once (for ncu): python tools/prof_eval.py [N]
  eval_kernel<3,2,0>        FP64 brute force, N x N, 3-D                     (C5 shape)
  eval_pruned_kernel<3,0>   the same through the error-bounded pruned route
  eval_f32_kernel<3,0>      FP32 variant
  eval_kernel<1,6,1>        LOO likelihood of a 1-D marginal, reference-order brute force (kdeb200_set_pruning(0))
  loo_sym_kernel<1,8>       the same through the symmetric each-pair-once kernel (default policy)   (C3 shape)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kde_b200 as K
import bench
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
K.init(0)
rng = np.random.default_rng(1)
pts, pos = bench.mixture(rng, 3, N), bench.mixture(rng, 3, N)
p = K.kde(pts, bench.silverman(pts))
for prec, name in ((K.F64, "f64 brute"), (K.F64_BOUNDED, "f64 bounded"), (K.F32, "f32")):
    K.evaluateDualTree(p, pos, precision=prec)
    print("eval", name, K.last_kernel_ms(), K.pruned_stats() if prec == K.F64_BOUNDED else "")
p1 = K.kde(pts[:1], [0.073])
for mode in (0, 1):
    K.set_pruning(mode)
    print("entropy mode", mode, K.entropy(p1), K.last_kernel_ms(), K.pruned_stats() if mode else "")
