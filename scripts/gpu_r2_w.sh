# round 2, run W: blend_normalize sweep (rows per warp x CTAs per SM)
mkdir -p gpurun_out
timeout 900 python scripts/tune_blend.py > gpurun_out/tune_blend.log 2>&1; echo "rc=$?"
cat gpurun_out/tune_blend.log | cut -c1-200
