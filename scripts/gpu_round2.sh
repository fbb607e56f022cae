#!/bin/bash
mkdir -p gpurun_out
for k in describe_kernel orient_kernel orient_exact_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:^$k -c 1 -f -o gpurun_out/full_$k python scripts/profile_step.py 512 1 > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_topk_kernel|rerank_kernel|cvt_f16" -c 6 -f -o gpurun_out/full_match_tc python scripts/profile_match.py 50000 > gpurun_out/ncu_match.log 2>&1; echo "ncu match rc=$?"
timeout 300 python scripts/profile_match.py 100000 > gpurun_out/match_100k.log 2>&1; tail -2 gpurun_out/match_100k.log
