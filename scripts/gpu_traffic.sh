# DRAM traffic of the filter kernel per step (ncu, dram__bytes_read/write per launch) at B = 4096, 64, 1:
# the "traffic" figure bench.py reports next to the algorithmic bytes.  --steps 2 --warmup 1, all launches kept.
mkdir -p gpurun_out
for B in 4096 64 1; do
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      -k "regex:score_filter_tc|spill_extract" --csv --log-file gpurun_out/traffic_b$B.csv \
      python bench.py --batch $B --steps 2 --warmup 1 --sweep "" --no-cpu-baseline --no-extras > gpurun_out/ncu_traffic_b$B.log 2>&1
  echo "traffic B=$B rc=$?"
done
