set -x
cd $GRAFT_REPO_ROOT
S=/usr/local/cuda/bin/compute-sanitizer
out=gpurun_out/sanitizer.txt
echo "compute-sanitizer on the B200 box (round 1, final kernels)" > $out
echo "== memcheck: smoke()" >> $out
timeout 300 $S --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^$" | tail -3 >> $out
echo "== racecheck: smoke()" >> $out
timeout 300 $S --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 >> $out
echo "== memcheck: native LCV (fused + host loop), eval-only handles" >> $out
timeout 400 $S --tool memcheck python -m pytest tests/test_gpu_eval.py -m gpu -q -x -k "native_lcv or eval_only or loo_eval" 2>&1 | tail -4 >> $out
echo "== racecheck: fused golden-section kernel" >> $out
timeout 400 $S --tool racecheck python -m pytest tests/test_gpu_eval.py -m gpu -q -x -k "native_lcv_equals_stepwise_mirror and (100 or 257)" 2>&1 | tail -4 >> $out
echo "== memcheck: Gibbs dynamic scheduling, tiny schedules" >> $out
timeout 400 $S --tool memcheck python -m pytest tests/test_gpu_gibbs.py -m gpu -q -x -k "many_batches or mixed_sizes" 2>&1 | tail -4 >> $out
echo "== racecheck: Gibbs dynamic scheduling" >> $out
timeout 400 $S --tool racecheck python -m pytest tests/test_gpu_gibbs.py -m gpu -q -x -k "many_batches and 129" 2>&1 | tail -4 >> $out
cat $out
