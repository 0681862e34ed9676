# Final validation of the round on one B200: GPU suite, smoke, both bench arms, launch lists, DRAM traffic, full captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_gpu.log
tail -14 gpurun_out/t_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?" >> gpurun_out/bench_ref.log
tail -2 gpurun_out/bench_ref.log
bash scripts/gpu_list.sh
bash scripts/gpu_traffic.sh
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_filter_tc -s 12 -c 4 -f \
    -o gpurun_out/prof_b4096 python bench.py --batch 4096 --steps 1 --warmup 1 --sweep "" --no-cpu-baseline > gpurun_out/ncu_full_b4096.log 2>&1
echo "full B=4096 rc=$?"
ls -la gpurun_out
