#!/bin/bash
# Tensor-core matcher: parity tests of both candidate kernels, then timings (path 3 = single CTA, 4 = CTA pair).
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_match_tc.py -m gpu -x -q -s > gpurun_out/pytest_tc.log 2>&1; echo "tc pytest rc=$?"; tail -12 gpurun_out/pytest_tc.log
nvidia-smi > /dev/null || echo "GPU unresponsive"
for n in ${MATCH_SIZES:-20000 100000}; do for path in 3 4; do
  timeout -s KILL 300 python scripts/profile_match.py $n $path 2>&1 | tail -1
done; done
