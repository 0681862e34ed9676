# round 2, run Y: top-1000 selects on the CTA-per-query bisection kernel -- k = 1000 parity tests, C5 shard on one GPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "1000 or c5 or exact_matches_oracle or peer_exchange" > gpurun_out/t_gpu_k1000.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_gpu_k1000.log
tail -3 gpurun_out/t_gpu_k1000.log
timeout 900 python bench.py --workload c5 --items 62500000 --steps 5 --warmup 3 --sweep "" --no-cpu-baseline > gpurun_out/bench_c5_n1.log 2>&1; echo "c5 rc=$?"
grep '^{' gpurun_out/bench_c5_n1.log | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['stage_ms'], j['e2e']['value'], j['roofline']['frac'])"
