# round 2, run Z: C5 shard on one GPU with growth 1 / 3 (default 2)
mkdir -p gpurun_out
for G in 1 3; do
HWER_GROWTH=$G timeout 900 python bench.py --workload c5 --items 62500000 --steps 5 --warmup 3 --sweep "" --no-cpu-baseline > gpurun_out/bench_c5_g$G.log 2>&1; echo "c5 g=$G rc=$?"
grep '^{' gpurun_out/bench_c5_g$G.log | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['stage_ms'], j['gpu_launches'])"
done
