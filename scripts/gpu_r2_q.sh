# round 2, run Q: evidence for the judged numbers on the in-tree build -- ncu --set full of the filter kernel's last
# three rounds at B = 4096 and the last round at B = 64, the dense select, DRAM traffic per step at B = 4096 / 64 / 1,
# launch lists at B = 4096 / 64, API stage timings
mkdir -p gpurun_out
X="--sweep '' --no-cpu-baseline --no-extras"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_filter_tc -s 13 -c 3 -f \
    -o gpurun_out/prof_b4096 python bench.py --batch 4096 --steps 1 --warmup 1 --sweep "" --no-cpu-baseline --no-extras > gpurun_out/ncu_full_b4096.log 2>&1
echo "full B=4096 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_filter_tc -s 7 -c 1 -f \
    -o gpurun_out/prof_b64 python bench.py --batch 64 --steps 1 --warmup 1 --sweep "" --no-cpu-baseline --no-extras > gpurun_out/ncu_full_b64.log 2>&1
echo "full B=64 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:select_compact_bisect -s 2 -c 1 -f \
    -o gpurun_out/prof_select python bench.py --batch 4096 --steps 1 --warmup 1 --sweep "" --no-cpu-baseline --no-extras > gpurun_out/ncu_full_select.log 2>&1
echo "full select rc=$?"
bash scripts/gpu_traffic.sh
bash scripts/gpu_list.sh
timeout 600 python scripts/time_api.py > gpurun_out/time_api.log 2>&1; cat gpurun_out/time_api.log | tail -14
