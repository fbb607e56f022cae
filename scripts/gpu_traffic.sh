#!/bin/bash
# DRAM bytes of every blur launch of one 512^3 extraction (two metrics only: one replay pass, small report).
mkdir -p gpurun_out
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"blur_march_kernel|blur_xy_kernel|blur_x_kernel" -c 200 -f -o gpurun_out/traffic_blur python scripts/profile_step.py 512 1 > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
ls -la gpurun_out/traffic_blur.ncu-rep
