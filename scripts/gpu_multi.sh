#!/bin/bash
# N-GPU visit: sharded parity + timings, then both bench lines at N.   usage: gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py > gpurun_out/multi_check_n$N.json 2> gpurun_out/multi_check_n$N.err; echo "multi_check rc=$?"; tail -3 gpurun_out/multi_check_n$N.err; cat gpurun_out/multi_check_n$N.json
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_n$N.err
grep "^{" gpurun_out/bench_n$N.json > gpurun_out/bench_n$N.clean.json; python scripts/show_bench.py gpurun_out/bench_n$N.clean.json | head -3
