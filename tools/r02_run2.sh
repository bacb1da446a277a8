#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gibbs.py tests/test_gpu_random.py tests/test_gpu_multi.py tests/test_abi.py -m gpu -x -q 2>&1 | tail -6
M=smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
for v in "" _ck32 _ck48; do
  so=$PWD/kerneldensityestimate.jl_b200/libkdeb200$v.so
  echo "== variant '$v'"
  KDEB200_SO=$so timeout 300 python tools/prof_gibbs.py 1000000 3 2>&1 | tail -2
  KDEB200_SO=$so timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/issued_gibbs$v.csv -k regex:gibbs_kernel python tools/prof_gibbs.py 75776 1 > /dev/null 2>&1
  grep -E "dram__bytes|gpu__time|fp64_cycles" gpurun_out/issued_gibbs$v.csv | cut -d, -f13-15
done
timeout 200 python tests/perf/bench_small.py 2>/dev/null | tail -30
