#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02_pytest_gpu.log
tail -6 gpurun_out/r02_pytest_gpu.log
timeout 600 python bench.py --steps 2 --no-cpu-baseline > gpurun_out/r02_bench_v2.json 2> gpurun_out/r02_bench_v2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_v2.json'))
print(d['value'], d['e2e']['value'])
s=d['secondary']
print(json.dumps(s['c5']['f64_bounded'],indent=1))
print(json.dumps(s['c3'],indent=1)[:1800])
PY
tail -3 gpurun_out/r02_bench_v2.err
