# round 2, run AS: small final kernel with warp-prefetched keys, 5 or 6 CTAs per SM, vs HEAD
mkdir -p gpurun_out
timeout 900 python scripts/ab_rounds.py variants/libhwer_b200_head.so variants/libhwer_b200_fin5.so variants/libhwer_b200_fin6.so > gpurun_out/ab_rounds.log 2>&1; echo "ab rc=$?"
grep "B=4096\|FAILED" gpurun_out/ab_rounds.log | sed 's/rounds \[.*\] chk/chk/' | cut -c1-220
