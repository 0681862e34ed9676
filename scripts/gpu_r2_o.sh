# round 2, run O: GPU suite, NCF bench (tensor maps encoded once), default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_gpu.log
tail -3 gpurun_out/t_gpu.log
timeout 300 python scripts/bench_ncf.py > gpurun_out/bench_ncf.log 2>&1; echo "ncf rc=$?"
tail -1 gpurun_out/bench_ncf.log | cut -c1-500
timeout 900 python bench.py --steps 20 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
grep '^{' gpurun_out/bench.log | cut -c1-400
