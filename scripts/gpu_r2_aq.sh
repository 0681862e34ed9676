# round 2, run AQ: chunked batch call with overlapped result copies -- parity test, C3 line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_gpu.log
tail -4 gpurun_out/t_gpu.log
timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.log 2>&1; echo "c3 rc=$?"
grep '^{' gpurun_out/bench_c3.log | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['e2e'])"
tail -5 gpurun_out/bench_c3.log | grep -i "error\|Traceback" | head
