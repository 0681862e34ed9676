# round 2, run M: bisection select (keys in registers) vs the radix select, same box; then the whole GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_gpu.log
tail -4 gpurun_out/t_gpu.log
timeout 900 python scripts/ab_rounds.py variants/libhwer_b200_hits.so variants/libhwer_b200_bisect.so > gpurun_out/ab_rounds.log 2>&1; echo "ab rc=$?"
cat gpurun_out/ab_rounds.log | cut -c1-250 | tail -14
timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 > gpurun_out/bench_c3.log 2>&1; echo "c3 rc=$?"
grep '^{' gpurun_out/bench_c3.log | cut -c1-1200
