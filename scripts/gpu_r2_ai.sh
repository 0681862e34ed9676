# round 2, run AI: epilogue with one 32-column chunk in registers at a time (no spills in the tile loop) vs HEAD
mkdir -p gpurun_out
timeout 900 python scripts/ab_rounds.py variants/libhwer_b200_head.so variants/libhwer_b200_onechunk.so > gpurun_out/ab_rounds.log 2>&1; echo "ab rc=$?"
cat gpurun_out/ab_rounds.log | cut -c1-250 | tail -14
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_gpu.log
tail -3 gpurun_out/t_gpu.log
