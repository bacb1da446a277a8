"""K1f (FP32 Gibbs sampler) against K1 (FP64) at the C4 shape: python tools/bench_gibbs_f32.py [Np] [d] -> one JSON object.
Kernel times are the library's own CUDA-event times; `same_labels` is the fraction of samples whose 8 labels agree when both
kernels are fed the same Philox streams; MUFU roofline: 2 MUFU per leaf-class pair (1 per node), 2 per internal sampleIndex
node (rsqrt + ex2), 1 per internal sampleIndices! node."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kde_b200 as K
from tests.util import mixture, silverman
Np = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 3
os.environ["KDEB200_GIBBS_WARP_MAX"] = "0"
K.init(0)
mufu, _ = K.pipe_peak(2, 200000)
rng = np.random.default_rng(2)
trees = []
for j in range(8):
    p = mixture(rng, d, 4096, 0.25 * j)
    trees.append(K.kde(p, silverman(p)))
out = {"samples": Np, "d": d, "mufu_lane_ops_per_s": mufu}
res = {}
for name, prec in (("f64", K.F64), ("f32", K.F32)):
    K.set_gibbs_precision(prec)
    ms = []
    for r in range(3):
        pts, idx = K.prodAppxMSGibbsS(None, trees, None, None, Niter=5, Np=Np, seed=1)
        ms.append(K.last_kernel_ms()[0])
    res[name] = (pts, idx)
    out[name] = {"kernel_ms": ms, "samples_per_s": Np / min(ms) * 1e3}
K.set_gibbs_precision(K.F64)
out["slow_draws"] = K.gibbs_f32_slow_draws()
same = np.all(res["f64"][1] == res["f32"][1], axis=0)
out["same_labels"] = float(same.mean())
out["max_point_diff_same_labels"] = float(np.max(np.abs(res["f64"][0][:, same] - res["f32"][0][:, same])))
out["mean_f64"] = res["f64"][0].mean(1).tolist()
out["mean_f32"] = res["f32"][0].mean(1).tolist()
out["std_f64"] = res["f64"][0].std(1).tolist()
out["std_f32"] = res["f32"][0].std(1).tolist()
L, pu, pn, ev = K.gibbs_sizes(trees, 5)
leaf = 8 * 2 * 4096 * 6
internal = ev - leaf
mufu_ops = leaf + internal * (5.0 / 6.0) * 2 + internal / 6.0
out["mufu_per_sample"] = mufu_ops
out["f32"]["mufu_roofline_frac"] = mufu_ops * out["f32"]["samples_per_s"] / mufu
out["speedup"] = out["f32"]["samples_per_s"] / out["f64"]["samples_per_s"]
print(json.dumps(out, indent=1))
