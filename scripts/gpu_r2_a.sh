#!/bin/bash
# round 2, visit A: the whole GPU suite after the stage refactor (incl. the new 512^3 / 10k / guard tests), a short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -30 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --no-cpu-baseline --match-n 0 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_a.err
python scripts/show_bench.py gpurun_out/bench_a.json
