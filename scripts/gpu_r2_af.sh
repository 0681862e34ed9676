# round 2, run AF: C3 (138,493 users x 27,278 items) under round-schedule variants
mkdir -p gpurun_out
for V in "0 0" "4096 6" "4096 3" "2048 4" "2048 13" "8192 3"; do
set -- $V
if [ "$1" = "0" ]; then unset HWER_FIRST_ROWS HWER_GROWTH; else export HWER_FIRST_ROWS=$1 HWER_GROWTH=$2; fi
timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_v.log 2>&1
grep '^{' gpurun_out/bench_c3_v.log | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('$V', round(j['value']), round(j['ms_per_step'],2), {k:round(v,2) for k,v in j['stage_ms'].items()}, j['result_checksum'])"
done
