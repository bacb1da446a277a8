#!/bin/bash
# final multi-GPU record of the sampler, one plain-C process driving all GPUs: tools/r02_multi_final.sh N  (gpurun --gpus N)
N=${1:-8}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
gcc -std=c99 -O2 -Iinclude examples/product_multi_c.c -o /tmp/product_multi_c -Lkerneldensityestimate.jl_b200 -lkdeb200 -lm -Wl,-rpath,$PWD/kerneldensityestimate.jl_b200
timeout 300 /tmp/product_multi_c 0 1000000 > gpurun_out/r02_final_multi_c_n${N}.json 2> gpurun_out/r02_final_multi_c_n${N}.err
KDEB200_GIBBS_F32=1 timeout 300 /tmp/product_multi_c 0 1000000 > gpurun_out/r02_final_multi_c_f32_n${N}.json 2> gpurun_out/r02_final_multi_c_f32_n${N}.err
cat gpurun_out/r02_final_multi_c_n${N}.json gpurun_out/r02_final_multi_c_f32_n${N}.json; tail -2 gpurun_out/r02_final_multi_c_f32_n${N}.err
