# Full ncu capture (with SASS source counters) of the hit-dense rounds 1..3 of the second B=4096 step.
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_filter_tc -s 9 -c 3 -f \
    -o gpurun_out/prof_early_b4096 python bench.py --batch 4096 --steps 1 --warmup 1 --sweep "" --no-cpu-baseline > gpurun_out/ncu_early.log 2>&1
echo "early rounds rc=$?"
ls -la gpurun_out/*.ncu-rep
