# round 2, run C: fused extract, warp-per-query final, A/B of each feature
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_gpu.log
tail -12 gpurun_out/t_gpu.log
timeout 900 python scripts/ab_knobs.py --variants 0,8,4,16,1 > gpurun_out/ab.log 2>&1; echo "ab rc=$?" >> gpurun_out/ab.log
grep -v RESULT gpurun_out/ab.log | tail -32
timeout 600 python bench.py --steps 20 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log | cut -c1-1200
bash scripts/gpu_list.sh
