"""FP32 (K1f) against FP64 (K1) Gibbs kernel per dimension at the C4 shape: python tools/bench_gibbs_f32_dims.py [Np] -> JSON."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kde_b200 as K
from tests.util import mixture, silverman
Np = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
os.environ["KDEB200_GIBBS_WARP_MAX"] = "0"
K.init(0)
rng = np.random.default_rng(2)
out = {"samples": Np, "dims": {}}
for d in range(1, 9):
    trees = []
    for j in range(8):
        p = mixture(rng, d, 4096, 0.25 * j)
        trees.append(K.kde(p, silverman(p)))
    r = {}
    idx = {}
    for name, prec in (("f64", K.F64), ("f32", K.F32)):
        K.set_gibbs_precision(prec)
        ms = []
        for rep in range(2):
            pts, ii = K.prodAppxMSGibbsS(None, trees, None, None, Niter=5, Np=Np, seed=1)
            ms.append(K.last_kernel_ms()[0])
        idx[name] = ii
        r[name + "_ms"] = min(ms)
        r[name + "_samples_per_s"] = Np / min(ms) * 1e3
    K.set_gibbs_precision(K.F64)
    r["speedup"] = r["f64_ms"] / r["f32_ms"]
    r["same_labels"] = float(np.all(idx["f64"] == idx["f32"], axis=0).mean())
    r["slow_draws"] = K.gibbs_f32_slow_draws()
    out["dims"][d] = r
print(json.dumps(out, indent=1))
