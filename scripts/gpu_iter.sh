# One kernel iteration on a B200: GPU suite, bench line, launch list of one step, schedule sweep at the headline batch.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_gpu.log
tail -4 gpurun_out/t_gpu.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log
K='regex:score_filter|select_compact|final_kernel|query_margin|fill_f32|merge_kernel'
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 120 --csv \
    --log-file gpurun_out/launches_b4096.csv python bench.py --steps 2 --warmup 1 --sweep "" --no-cpu-baseline \
    > gpurun_out/ncu_bench_b4096.log 2>&1
echo "launch list rc=$?"
if [ -n "$TUNE_BATCHES" ]; then timeout 300 python scripts/tune_schedule.py --batches $TUNE_BATCHES --steps 10 > gpurun_out/tune.log 2>&1; fi
[ -n "$TUNE_BATCHES" ] && tail -2 gpurun_out/tune.log
