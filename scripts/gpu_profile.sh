# Full ncu captures (with SASS source counters) of the dominant kernel: the last three rounds of a B=4096 step
# (tensor-bound) and every round of a B=64 step (HBM-bound); NCF re-rank throughput.  Numbers printed by bench.py
# under ncu are NOT bench values.
mkdir -p gpurun_out
timeout 600 python scripts/bench_ncf.py > gpurun_out/bench_ncf.log 2>&1; echo "ncf rc=$?"; tail -1 gpurun_out/bench_ncf.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_filter_tc -s 15 -c 3 -f \
    -o gpurun_out/prof_b4096 python bench.py --batch 4096 --steps 1 --warmup 1 --sweep "" --no-cpu-baseline > gpurun_out/ncu_full_b4096.log 2>&1
echo "full B=4096 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_filter_tc -s 5 -c 5 -f \
    -o gpurun_out/prof_b64 python bench.py --batch 64 --steps 1 --warmup 1 --sweep "" --no-cpu-baseline > gpurun_out/ncu_full_b64.log 2>&1
echo "full B=64 rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:ncf_linear -s 3 -c 3 -f \
    -o gpurun_out/prof_ncf python scripts/bench_ncf.py > gpurun_out/ncu_ncf.log 2>&1
echo "full ncf rc=$?"
