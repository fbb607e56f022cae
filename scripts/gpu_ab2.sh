#!/bin/bash
# A/B of the descriptor launch order + full ncu captures of octave 0's blur launches and the sparse kernels
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_sparse.py -m gpu -q -x > gpurun_out/pytest_ab.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_ab.log
for v in "S3D_DESC_ORDER=0" "S3D_DESC_ORDER=1"; do
  env $v timeout 300 python scripts/ab_step.py 512 4 2>&1 | tail -1
done | tee gpurun_out/ab2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'blur_march_kernel|blur_xyc_kernel|blur_xy_kernel' -c 12 -f -o gpurun_out/full_blur python scripts/profile_step.py 512 1 > gpurun_out/ncu_blur.log 2>&1; echo "ncu blur rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'describe_kernel|orient_kernel|orient_exact_kernel' -c 3 -f -o gpurun_out/full_sparse python scripts/profile_step.py 512 1 > gpurun_out/ncu_sparse.log 2>&1; echo "ncu sparse rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'detect_kernel|scan_kernel' -c 4 -f -o gpurun_out/full_detect python scripts/profile_step.py 512 1 > gpurun_out/ncu_detect.log 2>&1; echo "ncu detect rc=$?"
ls -la gpurun_out/*.ncu-rep
