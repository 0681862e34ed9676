# round 2, run AH: ncu --set full of a hit-dense mid round (round 3: 73,728 rows, 64 tiles per CTA) at B = 4096
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_filter_tc -s 11 -c 1 -f \
    -o gpurun_out/prof_round3 python bench.py --batch 4096 --steps 1 --warmup 1 --sweep "" --no-cpu-baseline --no-extras > gpurun_out/ncu_full_round3.log 2>&1
echo "full round3 rc=$?"
