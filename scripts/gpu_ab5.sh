#!/bin/bash
mkdir -p gpurun_out
S3D_ORIENT_VAR=2 timeout -s KILL 900 python -m pytest tests/test_gpu_sparse.py -m gpu -q -x > gpurun_out/pytest_ab.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_ab.log
for v in "S3D_ORIENT_VAR=1" "S3D_ORIENT_VAR=2" "S3D_ORIENT_VAR=3"; do
  env $v timeout 300 python scripts/ab_step.py 512 4 2>&1 | tail -1
done | tee gpurun_out/ab5.log
