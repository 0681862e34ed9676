# Full ncu capture (with SASS source counters) of rounds 3..7 of the second B=4096 step: round 3 is the last
# hit-heavy round, round 7 the longest tensor-bound one.  Numbers printed under ncu are not bench values.
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:score_filter_tc -s 11 -c 5 -f \
    -o gpurun_out/prof_rounds_b4096 python bench.py --batch 4096 --steps 1 --warmup 1 --sweep "" --no-cpu-baseline > gpurun_out/ncu_rounds.log 2>&1
echo "rounds rc=$?"
