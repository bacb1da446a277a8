#!/bin/bash
# Trimmed evidence refresh after an arithmetic-only kernel change (no sanitizer / launch list / small configs):
# tools/r02_evidence_trim.sh   (outputs under gpurun_out/)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r02_pytest_gpu.log
tail -3 gpurun_out/r02_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
M=smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
KDEB200_GIBBS_WARP_MAX=0 timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/issued_gibbs.csv -k regex:gibbs_kernel python tools/prof_gibbs.py 75776 1 > gpurun_out/issued_gibbs.log 2>&1
KDEB200_GIBBS_F32=1 KDEB200_GIBBS_WARP_MAX=0 timeout 600 ncu --metrics $M,smsp__inst_executed_pipe_xu.sum --clock-control none --csv --log-file gpurun_out/issued_gibbs_f32.csv -k regex:gibbs_f32_kernel python tools/prof_gibbs.py 303104 1 > gpurun_out/issued_gibbs_f32.log 2>&1
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/issued_eval.csv -k "regex:eval_kernel|eval_pruned_kernel|loo_sym_kernel" python tools/prof_eval.py 200000 > gpurun_out/issued_eval.log 2>&1
python tools/ncu_issued.py gpurun_out/ncu_issued.json "gibbs_c4:gibbs:gpurun_out/issued_gibbs.csv:gibbs_kernel<3:75776:sample" \
  "gibbs_f32_c4:gibbs_f32:gpurun_out/issued_gibbs_f32.csv:gibbs_f32_kernel<3:303104:sample" \
  "eval_c5:eval:gpurun_out/issued_eval.csv:eval_kernel<3:4e10:eval" "eval_c3:eval:gpurun_out/issued_eval.csv:eval_kernel<1:4e10:eval" \
  "eval_pruned_c5:eval_pruned:gpurun_out/issued_eval.csv:eval_pruned_kernel<3:4e10:eval" \
  "loo_sym_c3:eval_pruned:gpurun_out/issued_eval.csv:loo_sym_kernel<1:4e10:eval" > gpurun_out/ncu_issued.log 2>&1 || tail -5 gpurun_out/ncu_issued.log
mkdir -p profiles; cp gpurun_out/ncu_issued.json profiles/ncu_issued.json
KDEB200_GIBBS_WARP_MAX=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gibbs_kernel -c 1 -f -o gpurun_out/r02_gibbs python tools/prof_gibbs.py 75776 1 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r02_gibbs.ncu-rep gpurun_out/r02_gibbs_ncu.txt 2>/dev/null
timeout 400 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
timeout 600 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
timeout 300 python tools/bench_gibbs_f32.py 1000000 > gpurun_out/r02_gibbs_f32.json 2> /dev/null
cut -c1-300 gpurun_out/r02_bench.json; tail -2 gpurun_out/r02_bench.err; cut -c1-160 gpurun_out/r02_bench_reference.json
