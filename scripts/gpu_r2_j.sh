# round 2, run J: same-box A/B of filter-kernel builds (base = 96 registers with spills, regs112 = __maxnreg__(112),
# hits = regs112 + branch-free hit path), then the parity tests on the in-tree (= hits) build.
mkdir -p gpurun_out
timeout 900 python scripts/ab_rounds.py variants/libhwer_b200_base.so variants/libhwer_b200_hits.so > gpurun_out/ab_rounds.log 2>&1; echo "ab rc=$?"
cat gpurun_out/ab_rounds.log | cut -c1-600 | tail -30
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_gpu.log
tail -4 gpurun_out/t_gpu.log
