# round 2, run E: ncu --set full of the filter kernel's last three rounds at B = 4096 (scaled-query build), B = 64
# last round, DRAM traffic per step at B = 4096 / 64 / 1
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_filter_tc -s 13 -c 3 -f \
    -o gpurun_out/prof_b4096 python bench.py --batch 4096 --steps 1 --warmup 1 --sweep "" --no-cpu-baseline --no-extras > gpurun_out/ncu_full_b4096.log 2>&1
echo "full B=4096 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_filter_tc -s 7 -c 1 -f \
    -o gpurun_out/prof_b64 python bench.py --batch 64 --steps 1 --warmup 1 --sweep "" --no-cpu-baseline --no-extras > gpurun_out/ncu_full_b64.log 2>&1
echo "full B=64 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:final_kernel -s 1 -c 1 -f \
    -o gpurun_out/prof_final python bench.py --batch 4096 --steps 1 --warmup 1 --sweep "" --no-cpu-baseline --no-extras > gpurun_out/ncu_full_final.log 2>&1
echo "full final rc=$?"
sed -i 's/--no-cpu-baseline >/--no-cpu-baseline --no-extras >/' scripts/gpu_traffic.sh
bash scripts/gpu_traffic.sh
timeout 600 python scripts/time_api.py > gpurun_out/time_api.log 2>&1; cat gpurun_out/time_api.log | tail -14
