#!/bin/bash
# Short end-of-session visit: every GPU parity test, smoke, both bench arms (no ncu).
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python scripts/show_bench.py gpurun_out/bench.json | head -2
python -c "
import json;d=json.load(open('gpurun_out/bench.json'));print(d['e2e']['host_wall_ms_per_step'], d['roofline']['frac'], d['match']['ms'])"
