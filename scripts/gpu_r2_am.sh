# round 2, run AM: re-test of the feature knobs on the one-chunk epilogue: 8 = spill extraction fused into the filter kernel
mkdir -p gpurun_out
timeout 900 python scripts/ab_knobs.py --variants 0,8 --batches 4096,64 --reps 3 > gpurun_out/ab.log 2>&1; echo "ab rc=$?"
grep -v RESULT gpurun_out/ab.log | cut -c1-300
