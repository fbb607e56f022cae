#!/bin/bash
# parity tests for the dense/sparse/slab paths, then A/B of the run-time switches
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_sparse.py tests/test_gpu_slab.py -m gpu -q -x > gpurun_out/pytest_ab.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_ab.log
S3D_BLUR_XY=3 timeout -s KILL 600 python -m pytest tests/test_gpu_dense.py -m gpu -q -x > gpurun_out/pytest_ab3.log 2>&1; echo "pytest xy3 rc=$?"; tail -2 gpurun_out/pytest_ab3.log
S3D_ZVAR=2 timeout -s KILL 600 python -m pytest tests/test_gpu_dense.py -m gpu -q -x > gpurun_out/pytest_abz.log 2>&1; echo "pytest z2 rc=$?"; tail -2 gpurun_out/pytest_abz.log
for v in "S3D_BLUR_XY=1 S3D_ZVAR=0 S3D_DESC_ORDER=0" "S3D_BLUR_XY=2 S3D_ZVAR=0 S3D_DESC_ORDER=0" "S3D_BLUR_XY=3 S3D_ZVAR=0 S3D_DESC_ORDER=0" \
         "S3D_BLUR_XY=2 S3D_ZVAR=1 S3D_DESC_ORDER=0" "S3D_BLUR_XY=2 S3D_ZVAR=2 S3D_DESC_ORDER=0" "S3D_BLUR_XY=2 S3D_ZVAR=1 S3D_DESC_ORDER=1"; do
  env $v timeout 300 python scripts/ab_step.py 512 4 2>&1 | tail -1
done | tee gpurun_out/ab1.log
