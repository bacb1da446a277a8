"""Secondary workloads of BASELINE.json (parity-test cases with timings, not the bench line):
  C5: brute-force evaluation of a 1M-component 3-D KDE at 1M query points (FP64 and FP32)
  C3: kde! LOOCV bandwidth selection on 100k synthetic 4-D mixture points (FP64)
usage: python tests/perf/bench_eval.py [c5] [c3] [--small]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import kde_b200 as K
from tests.util import mixture, silverman

small = "--small" in sys.argv
which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c5", "c3"]
K.init(0)
dfma, _ = K.pipe_peak(0, 100000)
mufu, _ = K.pipe_peak(2, 100000)
out = {"dfma_lane_ops_per_s": dfma, "mufu_lane_ops_per_s": mufu}
if "c5" in which:
    N = M = 100_000 if small else 1_000_000
    rng = np.random.default_rng(20261017)
    pts, pos = mixture(rng, 3, N), mixture(rng, 3, M)
    t0 = time.perf_counter(); p = K.kde(pts, silverman(pts)); t_build = time.perf_counter() - t0
    for prec, name, slots in ((K.F64, "f64", 21), (K.F32, "f32", 1)):
        K.evaluateDualTree(p, pos[:, :1000], precision=prec)
        t0 = time.perf_counter(); v = K.evaluateDualTree(p, pos, precision=prec); wall = time.perf_counter() - t0
        ms, nl = K.last_kernel_ms()
        evals = float(N) * M
        peak = dfma if prec == K.F64 else mufu
        out["c5_" + name] = {"evals": evals, "kernel_ms": ms, "evals_per_s": evals / (ms * 1e-3), "e2e_s": wall,
                              "algorithmic_slots_per_eval": slots, "roofline_frac": evals * slots / (ms * 1e-3) / peak,
                              "bound": "fp64 pipe (2d+1+14 slots)" if prec == K.F64 else "MUFU ex2 (1/eval)", "launches": nl}
        if prec == K.F64:
            ref = v
        else:
            out["c5_f32"]["max_rel_err_vs_f64"] = float(np.max(np.abs(v - ref) / ref))
    out["c5_tree_build_host_s"] = t_build
    if "--no-cpu" not in sys.argv:  # CPU port on a bounded sample of the queries
        from oracle import oracle as O
        import bench as _b
        cores = _b.host_cores()
        o = O.OKDE.kde_bw(pts, silverman(pts))
        mq = 256
        t0 = time.perf_counter(); o.evaluate(pos[:, :mq]); t1 = time.perf_counter() - t0
        mq2 = 256 * cores
        t0 = time.perf_counter(); o.evaluate(pos[:, :mq2], nthreads=cores); tn = time.perf_counter() - t0
        out["c5_cpu_port"] = {"one_thread_evals_per_s": float(N) * mq / t1, "all_cores_evals_per_s": float(N) * mq2 / tn, "cores": cores,
                              "sample": "%d and %d of %d queries against all %d components" % (mq, mq2, M, N)}
if "c3" in which:
    N = 20_000 if small else 100_000
    rng = np.random.default_rng(3)
    pts = mixture(rng, 4, N)
    p1 = K.marginal(K.kde(pts, [1.0]), [1])
    K.entropy(p1)
    t0 = time.perf_counter(); H = K.entropy(p1); one = time.perf_counter() - t0
    ms, nl = K.last_kernel_ms()
    evals = float(N) * N
    out["c3_one_nLOO_LL"] = {"N": N, "kernel_ms": ms, "wall_ms": one * 1e3, "evals_per_s": evals / (ms * 1e-3),
                             "roofline_frac": evals * 17 / (ms * 1e-3) / dfma, "algorithmic_slots_per_eval": 17, "H": H, "launches": nl}
    if "--no-cpu" not in sys.argv:
        from oracle import oracle as O
        import bench as _b
        cores = _b.host_cores()
        o1 = O.OKDE.kde_bw(K.getPoints(p1), K.getBW(p1)[:, 0], K.getWeights(p1))
        mq = 64 * cores
        t0 = time.perf_counter(); o1.evaluate(K.getPoints(p1)[:, :mq], nthreads=cores); tn = time.perf_counter() - t0
        out["c3_cpu_port"] = {"all_cores_evals_per_s": float(N) * mq / tn, "cores": cores,
                              "sample": "%d of %d rows of one nLOO_LL (same arithmetic, no self-skip)" % (mq, N),
                              "extrapolated_one_nLOO_LL_s": tn * N / mq}
    t0 = time.perf_counter(); p = K.kde(pts); total = time.perf_counter() - t0
    out["c3_full_kde_lcv"] = {"N": N, "dims": 4, "wall_s": total, "bandwidth": K.getBW(p)[:, 0].tolist()}
print(json.dumps(out, indent=1))
