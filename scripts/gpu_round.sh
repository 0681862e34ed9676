# One full validation pass on a B200: GPU parity suite, smoke, bench (both arms), ncu launch lists, full captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t2.log 2>&1; echo "t2 rc=$?" >> gpurun_out/t2.log
tail -15 gpurun_out/t2.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -3 gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?" >> gpurun_out/bench_ref.log
tail -3 gpurun_out/bench_ref.log
K='regex:score_filter|select_compact|final_kernel|query_margin|fill_f32|blend_normalize|norm_stats|make_shadow|merge_kernel'
for B in 4096 64 1; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv \
      --log-file gpurun_out/launches_b$B.csv python bench.py --batch $B --steps 2 --warmup 1 --sweep "" --no-cpu-baseline \
      > gpurun_out/ncu_bench_b$B.log 2>&1
  echo "launch list B=$B rc=$?"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_filter_tc -s 6 -c 2 \
    -o gpurun_out/prof_b4096 python bench.py --batch 4096 --steps 1 --warmup 1 --sweep "" --no-cpu-baseline > gpurun_out/ncu_full_b4096.log 2>&1
echo "full B=4096 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_filter_tc -s 5 -c 2 \
    -o gpurun_out/prof_b64 python bench.py --batch 64 --steps 1 --warmup 1 --sweep "" --no-cpu-baseline > gpurun_out/ncu_full_b64.log 2>&1
echo "full B=64 rc=$?"
ls -la gpurun_out
