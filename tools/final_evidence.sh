#!/bin/bash
# Round-end evidence refresh on the GPU box: tools/final_evidence.sh  (outputs under gpurun_out/)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 400 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 python tests/perf/bench_eval.py > gpurun_out/bench_eval.json 2> gpurun_out/bench_eval.err
timeout 200 python tests/perf/bench_small.py > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -1; cut -c1-200 gpurun_out/bench.json; cut -c1-200 gpurun_out/bench_reference.json; wc -l gpurun_out/launches.csv
