# round 2, run AV: final kernel with both column blocks in flight at d = 256 -- parity tests, C3 line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_gpu.log
tail -3 gpurun_out/t_gpu.log
timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.log 2>&1; echo "c3 rc=$?"
grep '^{' gpurun_out/bench_c3.log | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['stage_ms'], j['e2e']['value'], j['result_checksum'])"
