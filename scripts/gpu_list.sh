# ncu launch lists (one step each) of the in-tree build at B = 4096 and 64.
mkdir -p gpurun_out
K='regex:score_filter|spill_extract|select_compact|final_kernel|query_margin|fill_f32|merge_kernel'
for B in 4096 64; do
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 90 --csv \
    --log-file gpurun_out/launches_b$B.csv python bench.py --batch $B --steps 2 --warmup 1 --sweep "" --no-cpu-baseline --no-extras \
    > gpurun_out/ncu_bench_b$B.log 2>&1
echo "launch list B=$B rc=$?"
done
