# round 2, run G: 148-SM work split, NCF on tcgen05, schedule sweep at B = 4096
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_gpu.log
tail -12 gpurun_out/t_gpu.log
timeout 300 python scripts/bench_ncf.py > gpurun_out/bench_ncf.log 2>&1; echo "ncf rc=$?" >> gpurun_out/bench_ncf.log
tail -2 gpurun_out/bench_ncf.log | cut -c1-800
timeout 600 python scripts/ab_knobs.py --variants 0,32 --batches 4096,512 > gpurun_out/ab.log 2>&1; echo "ab rc=$?" >> gpurun_out/ab.log
grep -v RESULT gpurun_out/ab.log | tail -10
timeout 900 python scripts/tune_schedule.py --batches 4096 --steps 10 > gpurun_out/tune.log 2>&1; echo "tune rc=$?" >> gpurun_out/tune.log
grep -v "^{" gpurun_out/tune.log | tail -24
timeout 900 python bench.py --steps 20 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log | cut -c1-1500
