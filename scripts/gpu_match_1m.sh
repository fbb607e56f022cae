#!/bin/bash
# 1M x 1M enhancedMatch on both tensor-core kernels, clocks sampled during the runs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv,noheader -lms 100 > gpurun_out/clocks_match.csv &
SMI=$!
for path in ${PATHS:-3 4}; do
  echo "path $path"; timeout -s KILL 600 python scripts/profile_match.py ${1:-1000000} $path 2>&1 | tail -2
done
kill $SMI
sort gpurun_out/clocks_match.csv | uniq -c | sort -rn | head -8
