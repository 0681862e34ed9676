# Validation pass of the final round-1 state on one B200: GPU suite, smoke, both bench arms, launch list, schedule sweep.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_gpu.log
tail -25 gpurun_out/t_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?" >> gpurun_out/bench_ref.log
tail -2 gpurun_out/bench_ref.log
K='regex:score_filter|select_compact|final_kernel|query_margin|fill_f32|blend_normalize|norm_stats|make_shadow|merge_kernel'
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv \
    --log-file gpurun_out/launches_b4096.csv python bench.py --steps 2 --warmup 1 --sweep "" --no-cpu-baseline \
    > gpurun_out/ncu_bench_b4096.log 2>&1
echo "launch list rc=$?"
timeout 300 python scripts/tune_schedule.py --batches 1,64 > gpurun_out/tune.log 2>&1; echo "tune rc=$?" >> gpurun_out/tune.log
tail -3 gpurun_out/tune.log
ls -la gpurun_out
