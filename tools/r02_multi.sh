#!/bin/bash
# multi-GPU evidence: tools/r02_multi.sh N   (run with gpurun --gpus N)
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_multi_n${N}_smi.txt
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_dist.py -q 2>&1 | tail -8 > gpurun_out/r02_multi_n${N}_pytest.log
tail -4 gpurun_out/r02_multi_n${N}_pytest.log
gcc -std=c99 -O2 -Iinclude examples/product_multi_c.c -o /tmp/product_multi_c -Lkerneldensityestimate.jl_b200 -lkdeb200 -lm -Wl,-rpath,$PWD/kerneldensityestimate.jl_b200
timeout 600 /tmp/product_multi_c 0 1000000 > gpurun_out/r02_product_multi_c_n${N}.json 2> gpurun_out/r02_product_multi_c_n${N}.err
cat gpurun_out/r02_product_multi_c_n${N}.json; tail -2 gpurun_out/r02_product_multi_c_n${N}.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_n${N}.json 2> gpurun_out/r02_bench_n${N}.err
cut -c1-300 gpurun_out/r02_bench_n${N}.json; tail -3 gpurun_out/r02_bench_n${N}.err
timeout 300 python tools/bench_multi_inproc.py > gpurun_out/r02_inproc_n${N}.json 2> gpurun_out/r02_inproc_n${N}.err
cat gpurun_out/r02_inproc_n${N}.json; tail -2 gpurun_out/r02_inproc_n${N}.err
