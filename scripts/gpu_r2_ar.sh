# round 2, run AR: ncu --set full of the small-list final kernel (first final launch of the second step) at B = 4096
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:final_kernel -s 2 -c 1 -f \
    -o gpurun_out/prof_final_small python bench.py --batch 4096 --steps 1 --warmup 1 --sweep "" --no-cpu-baseline --no-extras > gpurun_out/ncu_full_final_small.log 2>&1
echo "rc=$?"
