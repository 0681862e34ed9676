# Full ncu capture (with SASS source counters) of the mid-density rounds 4..5 of the second B=4096 step.
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_filter_tc -s 12 -c 2 -f \
    -o gpurun_out/prof_mid_b4096 python bench.py --batch 4096 --steps 1 --warmup 1 --sweep "" --no-cpu-baseline > gpurun_out/ncu_mid.log 2>&1
echo "mid rounds rc=$?"
