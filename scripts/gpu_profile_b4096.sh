mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_filter_tc -s 4 -c 2 \
    -o gpurun_out/prof_b4096 -f python bench.py --batch 4096 --steps 1 --warmup 1 --sweep "" --no-cpu-baseline > gpurun_out/ncu_full_b4096.log 2>&1
echo "full B=4096 rc=$?"
