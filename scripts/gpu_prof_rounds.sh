# Parity suite + bench + a full ncu capture of rounds 2..7 of one B=4096 step (hit-bound early rounds and the
# tensor-bound late rounds).  Numbers printed under ncu are not bench values.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/t2.log 2>&1; echo "t2 rc=$?" >> gpurun_out/t2.log
tail -5 gpurun_out/t2.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -3 gpurun_out/bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_filter_tc -s 2 -c 6 \
    -o gpurun_out/prof_rounds_b4096 python bench.py --batch 4096 --steps 1 --warmup 1 --sweep "" --no-cpu-baseline > gpurun_out/ncu_rounds.log 2>&1
echo "rounds rc=$?"
