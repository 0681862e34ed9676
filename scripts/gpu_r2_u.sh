# round 2, run U: pre-selection in the dense select vs HEAD (same box), parity tests, C3 line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_gpu.log
tail -3 gpurun_out/t_gpu.log
timeout 900 python scripts/ab_rounds.py variants/libhwer_b200_head.so variants/libhwer_b200_presel.so variants/libhwer_b200_presel2.so > gpurun_out/ab_rounds.log 2>&1; echo "ab rc=$?"
cat gpurun_out/ab_rounds.log | cut -c1-250 | tail -14
timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.log 2>&1; echo "c3 rc=$?"
grep '^{' gpurun_out/bench_c3.log | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['stage_ms'], j['e2e']['value'])"
