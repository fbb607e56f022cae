#!/bin/bash
mkdir -p gpurun_out
for v in "S3D_BLUR_XY_SEG=0" "S3D_BLUR_XY_SEG=256" "S3D_BLUR_XY_SEG=512" "S3D_MARCH_SEG=128" "S3D_MARCH_SEG=32"; do
  env $v timeout 300 python scripts/ab_step.py 512 4 2>&1 | tail -1
done | tee gpurun_out/ab6.log
