# Quick validation of a kernel change: parity suite (stop at first failure), bench, per-round launch list at B=4096.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/t2.log 2>&1; echo "t2 rc=$?" >> gpurun_out/t2.log
tail -25 gpurun_out/t2.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -3 gpurun_out/bench.log
K='regex:score_filter|select_compact|final_kernel'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 120 --csv \
    --log-file gpurun_out/launches_b4096.csv python bench.py --batch 4096 --steps 2 --warmup 1 --sweep "" --no-cpu-baseline \
    > gpurun_out/ncu_bench_b4096.log 2>&1
echo "launch list rc=$?"
