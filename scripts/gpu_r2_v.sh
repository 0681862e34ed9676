# round 2, run V: ncu --set full of the dense select (pre-selection build)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:select_compact_bisect -s 2 -c 1 -f \
    -o gpurun_out/prof_select2 python bench.py --batch 4096 --steps 1 --warmup 1 --sweep "" --no-cpu-baseline --no-extras > gpurun_out/ncu_full_select2.log 2>&1
echo "full select rc=$?"
