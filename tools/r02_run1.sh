#!/bin/bash
# round-2 GPU call 1: tests, bench (both arms), issued-instruction captures
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu.log
tail -5 gpurun_out/r02_pytest_gpu.log
M=smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/issued_gibbs.csv -k regex:gibbs_kernel python tools/prof_gibbs.py 75776 1 > gpurun_out/issued_gibbs.log 2>&1
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/issued_eval.csv -k regex:eval_kernel python tools/prof_eval.py 200000 > gpurun_out/issued_eval.log 2>&1
python tools/ncu_issued.py gpurun_out/ncu_issued.json "gibbs_c4:gibbs:gpurun_out/issued_gibbs.csv:gibbs_kernel<3:75776:sample" "eval_c5:eval:gpurun_out/issued_eval.csv:eval_kernel<3:4e10:eval" "eval_c3:eval:gpurun_out/issued_eval.csv:eval_kernel<1:4e10:eval" > gpurun_out/ncu_issued.log 2>&1
mkdir -p profiles; cp gpurun_out/ncu_issued.json profiles/ncu_issued.json
timeout 600 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
cut -c1-600 gpurun_out/r02_bench.json; tail -3 gpurun_out/r02_bench.err
