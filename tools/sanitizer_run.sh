#!/bin/bash
# compute-sanitizer over the round-2 kernels: tools/sanitizer_run.sh   (writes gpurun_out/r02_sanitizer.txt)
cd "$(dirname "$0")/.."
S=/usr/local/cuda/bin/compute-sanitizer
out=gpurun_out/r02_sanitizer.txt
mkdir -p gpurun_out
echo "compute-sanitizer on the B200 box (round 2 kernels)" > $out
run() {  # label, tool, pytest selection...
  echo "== $1" >> $out
  local tool=$2; shift 2
  timeout 600 $S --tool $tool python -m pytest "$@" -m gpu -q -x 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|hazard" | tail -4 >> $out
}
echo "== memcheck: smoke()" >> $out
timeout 300 $S --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^$" | tail -3 >> $out
run "memcheck: pruned evaluation + symmetric LOO (boxes, mask, order, exact pass)" memcheck tests/test_gpu_pruned.py -k "far_queries or symmetric and (5000 or 9000) or bounded_eval_matches and 9999"
run "racecheck: symmetric LOO kernel (shared-memory column credits)" racecheck tests/test_gpu_pruned.py -k "symmetric and 5000"
run "memcheck: warp-per-chain + thread-per-chain Gibbs, tiny schedules, masks, label recording" memcheck tests/test_gpu_gibbs.py -k "many_batches or partial_dim_mask or label_recording or mixed_sizes"
run "racecheck: warp-per-chain Gibbs" racecheck tests/test_gpu_gibbs.py -k "warp and many_batches and 129"
run "memcheck: fused product (on-chip sort + ball-tree statistics + golden sections), sample, marginals" memcheck tests/test_gpu_extras.py -k "product_in_one_call and (100 or 300 or 512) or sample_with_injected or eval_marginals and 777"
run "racecheck: fused product kernel" racecheck tests/test_gpu_extras.py -k "product_in_one_call and 3-3-64"
run "memcheck: FP32 Gibbs sampler (pair records, prep kernel, FP64 redo path, masks)" memcheck tests/test_gpu_gibbs_f32.py -k "same_streams and (3-6-100 or 2-2-5001 or 1-3-2 or 8-2-150) or underflow or poisoned"
run "racecheck: FP32 Gibbs sampler (TMA ring)" racecheck tests/test_gpu_gibbs_f32.py -k "same_streams and 2-2-100-2000"
run "memcheck: in-process multi-GPU (contexts on one device), schedule cache" memcheck tests/test_gpu_multi.py -k "gibbs_sharded and oversubscribed or small_calls"
cat $out
