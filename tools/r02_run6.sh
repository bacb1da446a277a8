#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/test_gpu_gibbs.py tests/test_gpu_random.py tests/test_abi.py tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02_pytest_gibbs.log
tail -5 gpurun_out/r02_pytest_gibbs.log
timeout 300 python tools/gibbs_kernel_crossover.py > gpurun_out/r02_gibbs_crossover.json 2> gpurun_out/r02_gibbs_crossover.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_gibbs_crossover.json'))
for k,rows in d.items():
    print(k); [print("   ",r) for r in rows]
PY
tail -3 gpurun_out/r02_gibbs_crossover.err
for v in "" _symq6 _symq4; do
echo "== variant '$v'"
KDEB200_SO=$PWD/kerneldensityestimate.jl_b200/libkdeb200$v.so python - <<'PY' 2>&1 | tail -8
import time, numpy as np, sys
sys.path.insert(0,'.')
import kde_b200 as K, bench
K.init(0)
pts = bench.mixture(np.random.default_rng(3), 4, 100_000)
p1 = K.marginal(K.kde(pts, [1.0]), [1])
for h in (4.0, 0.073):
    q = K.kde(K.getPoints(p1), [h]); K.entropy(q)
    ms=[]
    for r in range(5):
        K.entropy(q); ms.append(K.last_kernel_ms()[0])
    print("h=%.3f sym kernel %.3f ms"%(h, float(np.median(ms))))
K.kde(pts[:, :5000])
t0=time.perf_counter(); pk=K.kde(pts); print("kde!",time.perf_counter()-t0)
PY
done
timeout 200 python tests/perf/bench_small.py 2>/dev/null > gpurun_out/r02_bench_small.json; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_small.json'))
for k,v in d.items(): print(k, {a:round(b,3) for a,b in v.items() if isinstance(b,float)})"
