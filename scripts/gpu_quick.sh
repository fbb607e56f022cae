#!/bin/bash
# Parity tests + one bench line (+ optional extra command given as $1)
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python scripts/show_bench.py gpurun_out/bench.json
if [ -n "$1" ]; then bash -c "$1"; fi
