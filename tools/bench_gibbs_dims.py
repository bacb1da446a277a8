"""Gibbs kernel throughput and algorithmic roofline fraction per dimension at the C4 shape (8 x 4096 components, Niter=5):
python tools/bench_gibbs_dims.py [Np]   -> one JSON object.
Algorithmic FP64-pipe slots: leaf-class evaluation 2d + 2 + 14, internal-class 9d + 20 (= 22 / 47 at d = 3)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kde_b200 as K
from tests.util import mixture, silverman
Np = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
K.init(0)
dfma, _ = K.pipe_peak(0, 200000)
rng = np.random.default_rng(2)
out = {"samples": Np, "dfma_lane_ops_per_s": dfma, "dims": {}}
for d in range(1, 9):
    trees = []
    for j in range(8):
        p = mixture(rng, d, 4096, 0.25 * j)
        trees.append(K.kde(p, silverman(p)))
    for r in range(2):
        K.prodAppxMSGibbsS(None, trees, None, None, Niter=5, Np=Np, seed=1)
    ms, _ = K.last_kernel_ms()
    L, pu, pn, ev = K.gibbs_sizes(trees, 5)
    leaf = 8 * 2 * 4096 * 6            # the two all-leaf levels of a 4096-leaf tree, 8 densities, 1 + Niter passes
    internal = ev - leaf
    slots = leaf * (2 * d + 16) + internal * (9 * d + 20)
    out["dims"][d] = {"kernel_ms": ms, "samples_per_s": Np / ms * 1e3, "evals_per_s": Np * ev / ms * 1e3,
                      "algorithmic_fp64_slots_per_sample": slots, "roofline_frac": slots * Np / (ms * 1e-3) / dfma}
print(json.dumps(out, indent=1))
