# round 2, run N: ncu launch list of one C3 pass (every kernel), and of the NCF re-rank bench
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c3.csv \
    python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_c3.log 2>&1; echo "c3 list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ncf -c 120 --csv --log-file gpurun_out/launches_ncf.csv \
    python scripts/bench_ncf.py > gpurun_out/ncu_bench_ncf.log 2>&1; echo "ncf list rc=$?"
python scripts/summarize_launches.py gpurun_out/launches_c3.csv | tail -40 | cut -c1-2500
python scripts/summarize_launches.py gpurun_out/launches_ncf.csv | tail -12 | cut -c1-1500
