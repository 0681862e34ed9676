# round 2, run D: full GPU suite on the new defaults, bench with API-level e2e + c3 / c2 blocks, launch lists
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_gpu.log
tail -12 gpurun_out/t_gpu.log
timeout 900 python bench.py --steps 20 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -3 gpurun_out/bench.log | cut -c1-6000
timeout 600 python bench.py --workload c3 --steps 5 > gpurun_out/bench_c3.log 2>&1; echo "c3 rc=$?" >> gpurun_out/bench_c3.log
tail -3 gpurun_out/bench_c3.log | cut -c1-3000
bash scripts/gpu_list.sh
