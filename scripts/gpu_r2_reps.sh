#!/bin/bash
# Round-2 first visit: bring full ncu reports (with source) of the top kernels back for offline analysis.
mkdir -p gpurun_out /tmp/ncu
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
date +%s > gpurun_out/t0.txt
cap() {  # name regex skip count
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o /tmp/ncu/$1 python scripts/profile_step.py 512 1 > gpurun_out/ncu_$1.log 2>&1
  echo "ncu $1 rc=$?"
  python scripts/ncu_summary.py /tmp/ncu/$1.ncu-rep --src 60 > gpurun_out/r2_$1_summary.txt 2>&1
  ls -la /tmp/ncu/$1.ncu-rep
  xz -T0 -3 -c /tmp/ncu/$1.ncu-rep > gpurun_out/$1.ncu-rep.xz
}
cap describe 'describe_kernel' 0 1
cap orient 'orient_kernel' 0 1
cap detect 'detect_kernel' 0 1
cap xyc8 'blur_xyc_kernel<8>' 0 1
cap xyc5 'blur_xyc_kernel<5>' 0 1
cap xy3 'blur_xy_kernel<3>' 0 1
date +%s > gpurun_out/t1.txt
du -sh gpurun_out
