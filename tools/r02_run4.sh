#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/test_gpu_pruned.py tests/test_gpu_eval.py tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02_pytest_pruned.log
tail -6 gpurun_out/r02_pytest_pruned.log
python - <<'PY' 2>&1 | tail -30
import time, numpy as np, sys
sys.path.insert(0,'.')
import kde_b200 as K, bench
K.init(0)
pts = bench.mixture(np.random.default_rng(3), 4, 100_000)
p1 = K.marginal(K.kde(pts, [1.0]), [1])
for h in (4.0, 1.0, 0.3, 0.073):
    q = K.kde(K.getPoints(p1), [h])
    for mode in (0, 1):
        K.set_pruning(mode); K.entropy(q)
        t0=time.perf_counter(); H=K.entropy(q); dt=time.perf_counter()-t0
        ms,nl = K.last_kernel_ms()
        print("h=%.3f mode %d: H=%.15g kernel %.3f ms wall %.3f ms launches %d"%(h,mode,H,ms,dt*1e3,nl), K.pruned_stats() if mode else "")
for mode in (0,1):
    K.set_pruning(mode); K.kde(pts[:, :5000])
    t0=time.perf_counter(); pk=K.kde(pts); print("kde! mode",mode,time.perf_counter()-t0, K.getBW(pk)[:,0])
rng=np.random.default_rng(bench.SEED); N=1_000_000
cp, pos = bench.mixture(rng,3,N), bench.mixture(rng,3,N)
p=K.kde(cp,bench.silverman(cp)); K.evaluateDualTree(p,pos[:,:4096])
for prec in (K.F64_BOUNDED, K.F64_BOUNDED):
    t0=time.perf_counter(); v=K.evaluateDualTree(p,pos,precision=prec); dt=time.perf_counter()-t0
    print("C5 bounded: wall %.1f ms kernel %.1f ms"%(dt*1e3,K.last_kernel_ms()[0]), K.pruned_stats())
PY
