# ncu launch lists (per-kernel device time, cold cache/serialised: compare SHARES) and one full capture of the
# dominant kernel per batch size.  Numbers printed by bench.py under ncu are NOT bench values.
mkdir -p gpurun_out
K='regex:score_filter|select_compact|final_kernel|query_margin|fill_f32|blend_normalize|norm_stats|make_shadow|merge_kernel'
for B in 4096 64 1; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv \
      --log-file gpurun_out/launches_b$B.csv python bench.py --batch $B --steps 2 --warmup 1 --sweep "" --no-cpu-baseline \
      > gpurun_out/ncu_bench_b$B.log 2>&1
  echo "launch list B=$B rc=$?"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_filter_tc -s 4 -c 2 \
    -o gpurun_out/prof_b4096 python bench.py --batch 4096 --steps 1 --warmup 1 --sweep "" --no-cpu-baseline > gpurun_out/ncu_full_b4096.log 2>&1
echo "full B=4096 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_filter_tc -s 3 -c 2 \
    -o gpurun_out/prof_b64 python bench.py --batch 64 --steps 1 --warmup 1 --sweep "" --no-cpu-baseline > gpurun_out/ncu_full_b64.log 2>&1
echo "full B=64 rc=$?"
ls -la gpurun_out
