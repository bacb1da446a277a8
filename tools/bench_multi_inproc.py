"""C5 and C3 through the in-process multi-GPU C-ABI (kdeb200_init_multi): one python process, all visible GPUs.
Prints one JSON line: wall seconds on 1 GPU and on all GPUs for the 1M x 1M evaluation and the 100k x 4-D kde! LOOCV."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kde_b200 as K
import bench

n5 = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
n3 = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
K.init(0)
rng = np.random.default_rng(bench.SEED)
pts, pos = bench.mixture(rng, 3, n5), bench.mixture(rng, 3, n5)
p = K.kde(pts, bench.silverman(pts))
x3 = bench.mixture(np.random.default_rng(3), 4, n3)
out = {"c5_n": n5, "c3_n": n3}
for label, G in (("one_gpu", 1), ("all_gpus", 0)):
    g = K.init_multi(G)
    K.evaluateDualTree(p, pos[:, :8192 * g])       # replicate the tree, warm every device
    t0 = time.perf_counter(); v = K.evaluateDualTree(p, pos); t5 = time.perf_counter() - t0
    ms5, _ = K.last_kernel_ms()
    K.lcv_bandwidths(x3)                           # warm every device (lazy kernel loading, tree replicas, pools)
    ts = []
    for rep in range(3):
        t0 = time.perf_counter(); bw = K.lcv_bandwidths(x3); ts.append(time.perf_counter() - t0)
    t3 = min(ts)
    out[label] = {"n_gpus": g, "c5_wall_s": t5, "c5_slowest_kernel_ms": ms5, "c5_evals_per_s": float(n5) * n5 / t5,
                  "c3_kde_wall_s": t3, "c3_bandwidth": bw.tolist(), "c5_checksum": float(v.sum())}
out["c5_speedup"] = out["one_gpu"]["c5_wall_s"] / out["all_gpus"]["c5_wall_s"]
out["c3_speedup"] = out["one_gpu"]["c3_kde_wall_s"] / out["all_gpus"]["c3_kde_wall_s"]
print(json.dumps(out))
