#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_slab.py tests/test_gpu_dense.py -m gpu -q -x > gpurun_out/pytest_ab.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_ab.log
timeout 300 python scripts/ab_step.py 512 4 2>&1 | tail -1 | tee gpurun_out/ab4.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
