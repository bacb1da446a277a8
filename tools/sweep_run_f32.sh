#!/bin/bash
# C4-shaped Gibbs call with every libkdeb200_*.so tuning variant of the FP32 kernel: tools/sweep_run_f32.sh [samples]
N=${1:-1000000}
cd "$(dirname "$0")/.."
export KDEB200_GIBBS_WARP_MAX=0
echo "== base f64"; timeout 300 python tools/prof_gibbs.py $N 3 | tail -2
echo "== base f32"; KDEB200_GIBBS_F32=1 timeout 300 python tools/prof_gibbs.py $N 3 | tail -2
for so in kerneldensityestimate.jl_b200/libkdeb200_*.so; do
  echo "== $so"; KDEB200_GIBBS_F32=1 KDEB200_SO=$PWD/$so timeout 300 python tools/prof_gibbs.py $N 3 | tail -2
done
