#!/bin/bash
N=${1:-8}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_n${N}.json 2> gpurun_out/r02_bench_n${N}.err
cut -c1-200 gpurun_out/r02_bench_n${N}.json; tail -3 gpurun_out/r02_bench_n${N}.err
timeout 300 python tools/bench_multi_inproc.py > gpurun_out/r02_inproc_n${N}.json 2> gpurun_out/r02_inproc_n${N}.err
cat gpurun_out/r02_inproc_n${N}.json; tail -2 gpurun_out/r02_inproc_n${N}.err
