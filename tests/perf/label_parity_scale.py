"""Label parity at scale: C4-shaped product, injected Philox streams, GPU vs the oracle
(OpenMP over chains).  Reports the number of samples whose label vector differs.
usage: python tests/perf/label_parity_scale.py [samples] [components (default 4096: the C4 shape)] [densities]
(KDEB200_GIBBS_WARP_MAX=0 / =1000000000 forces the thread- / warp-per-chain kernel; the warp kernel needs level lists
that fit its shared memory, i.e. <= ~2900 components -- above that the library uses the thread kernel anyway)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import kde_b200 as K
from oracle import oracle as O
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ncomp = int(sys.argv[2]) if len(sys.argv) > 2 else bench.NCOMP
ndens = int(sys.argv[3]) if len(sys.argv) > 3 else bench.NDENS
K.init(0)
pts = [bench.synth_points(j)[:, :ncomp] for j in range(ndens)]
kt = [K.kde(p, bench.silverman(p)) for p in pts]
ot = [O.OKDE.kde_bw(p, bench.silverman(p)) for p in pts]
L, perU, perN, evals = K.gibbs_sizes(kt, bench.NITER)
U, G = K.philox_streams(bench.SEED, n, perU, perN)
t0 = time.perf_counter(); gp, gi = K.prodAppxMSGibbsS(None, kt, None, None, Niter=bench.NITER, Np=n, randU=U, randN=G); tg = time.perf_counter() - t0
cores = bench.host_cores()
t0 = time.perf_counter(); ep, ei = O.gibbs(ot, n, bench.NITER, U, G, nthreads=cores); tc = time.perf_counter() - t0
bad = int(np.sum(np.any(gi != ei, axis=0)))
print(json.dumps({"samples": n, "components": ncomp, "densities": ndens, "kernel": os.environ.get("KDEB200_GIBBS_WARP_MAX", "policy"), "label_draws": n * perU, "kernel_evals": n * evals, "samples_with_label_mismatch": bad,
                  "max_abs_point_diff": float(np.max(np.abs(gp - ep))), "max_abs_point": float(np.max(np.abs(ep))),
                  "gpu_call_s": tg, "oracle_s": tc, "oracle_threads": cores}))
