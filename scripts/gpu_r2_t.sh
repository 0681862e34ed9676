# round 2, run T: vectorised query staging + two-deep extract prefetch vs HEAD, same box; parity tests on the new build
mkdir -p gpurun_out
timeout 900 python scripts/ab_rounds.py variants/libhwer_b200_head.so variants/libhwer_b200_stage.so > gpurun_out/ab_rounds.log 2>&1; echo "ab rc=$?"
cat gpurun_out/ab_rounds.log | cut -c1-250 | tail -14
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_gpu.log
tail -3 gpurun_out/t_gpu.log
