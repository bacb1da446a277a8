#!/bin/bash
# runs the C4-shaped Gibbs kernel with every tuning variant: tools/sweep_run.sh [samples]
N=${1:-1000000}
cd "$(dirname "$0")/.."
echo "== base"; timeout 300 python tools/prof_gibbs.py $N 2 | tail -1
for so in kerneldensityestimate.jl_b200/libkdeb200_*.so; do
  echo "== $so"; KDEB200_SO=$PWD/$so timeout 300 python tools/prof_gibbs.py $N 2 | tail -1
done
