"""BASELINE configs 1 and 2 (the reference's own CPU-sized cases): GPU wall time per call next to the
oracle port on one thread.  usage: python tests/perf/bench_small.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import kde_b200 as K
from oracle import oracle as O

K.init(0)
rng = np.random.default_rng(1)
out = {}
cases = {
    "c1_readme_product": ([rng.standard_normal((2, 100)), 2.0 + rng.standard_normal((2, 100))], 100),
    "c2_beta_rayleigh": ([rng.beta(1.0, 0.45, size=(1, 300)), rng.rayleigh(0.5, size=(1, 100)) - 0.5], 10000),
}
for name, (pts, Np) in cases.items():
    kt = [K.kde(p) for p in pts]  # warm-up (module load, pools)
    t0 = time.perf_counter(); kt = [K.kde(p, native_lcv=False) for p in pts]; t_lcv_step = time.perf_counter() - t0
    tl = []
    for r in range(5):
        t0 = time.perf_counter(); kt = [K.kde(p) for p in pts]; tl.append(time.perf_counter() - t0)
    t_lcv = float(np.median(tl))
    t0 = time.perf_counter(); [K.lcv_bandwidths(p) for p in pts]; t_bw = time.perf_counter() - t0
    lcv_kernel_ms, _ = K.last_kernel_ms()
    tp = []
    for r in range(5):
        t0 = time.perf_counter(); pq = K.prod(kt, seed=r) if hasattr(K, "prod") else None; tp.append(time.perf_counter() - t0)
    t0 = time.perf_counter(); ot = [O.OKDE.kde_lcv(p) for p in pts]; t_lcv_cpu = time.perf_counter() - t0
    K.prodAppxMSGibbsS(None, kt, None, None, Niter=5, Np=Np, seed=1)
    ts = []
    for r in range(5):
        t0 = time.perf_counter(); K.prodAppxMSGibbsS(None, kt, None, None, Niter=5, Np=Np, seed=r); ts.append(time.perf_counter() - t0)
    ms, _ = K.last_kernel_ms()
    nU, nN = O.prod_sizes(ot, Np, 5)
    U, G = rng.random(nU), rng.standard_normal(nN)
    t0 = time.perf_counter(); O.gibbs(ot, Np, 5, U, G); t_cpu = time.perf_counter() - t0
    L, pu, pn, ev = K.gibbs_sizes(kt, 5)
    out[name] = {"samples": Np, "evals_per_sample": ev, "gpu_call_ms": 1e3 * float(np.median(ts)), "gpu_kernel_ms": ms,
                 "cpu_port_1thread_ms": 1e3 * t_cpu, "kde_lcv_gpu_ms": 1e3 * t_lcv, "kde_lcv_gpu_stepwise_ms": 1e3 * t_lcv_step,
                 "lcv_bandwidths_only_ms": 1e3 * t_bw, "lcv_last_kernel_ms": lcv_kernel_ms,
                 "product_with_lcv_refit_ms": 1e3 * float(np.median(tp)), "kde_lcv_cpu_port_ms": 1e3 * t_lcv_cpu,
                 "note": "latency-bound on the GPU (few chains); reported as wall time, not roofline"}
print(json.dumps(out, indent=1))
