"""Label parity at scale: C4-shaped product, injected Philox streams, GPU vs the oracle
(OpenMP over chains).  Reports the number of samples whose label vector differs.
usage: python tests/perf/label_parity_scale.py [samples]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import kde_b200 as K
from oracle import oracle as O
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
K.init(0)
pts = [bench.synth_points(j) for j in range(bench.NDENS)]
kt = [K.kde(p, bench.silverman(p)) for p in pts]
ot = [O.OKDE.kde_bw(p, bench.silverman(p)) for p in pts]
L, perU, perN, evals = K.gibbs_sizes(kt, bench.NITER)
U, G = K.philox_streams(bench.SEED, n, perU, perN)
t0 = time.perf_counter(); gp, gi = K.prodAppxMSGibbsS(None, kt, None, None, Niter=bench.NITER, Np=n, randU=U, randN=G); tg = time.perf_counter() - t0
cores = bench.host_cores()
t0 = time.perf_counter(); ep, ei = O.gibbs(ot, n, bench.NITER, U, G, nthreads=cores); tc = time.perf_counter() - t0
bad = int(np.sum(np.any(gi != ei, axis=0)))
print(json.dumps({"samples": n, "label_draws": n * perU, "kernel_evals": n * evals, "samples_with_label_mismatch": bad,
                  "max_abs_point_diff": float(np.max(np.abs(gp - ep))), "max_abs_point": float(np.max(np.abs(ep))),
                  "gpu_call_s": tg, "oracle_s": tc, "oracle_threads": cores}))
