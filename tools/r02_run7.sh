#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/test_gpu_gibbs.py tests/test_gpu_random.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02_pytest_gibbs.log
tail -3 gpurun_out/r02_pytest_gibbs.log
timeout 300 python tools/gibbs_kernel_crossover.py > gpurun_out/r02_gibbs_crossover.json 2> gpurun_out/r02_gibbs_crossover.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_gibbs_crossover.json'))
for k,rows in d.items():
    print(k); [print("   ",r) for r in rows[:5]]
PY
timeout 200 python tests/perf/bench_small.py 2>/dev/null > gpurun_out/r02_bench_small.json; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_small.json'))
for k,v in d.items(): print(k, {a:round(b,3) for a,b in v.items() if isinstance(b,float)})"
