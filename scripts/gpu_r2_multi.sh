#!/bin/bash
# N-GPU visit: NCCL one-process-per-GPU parity + timing, single-process multi-device parity + timing
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
nvidia-smi topo -m 2>/dev/null | head -12 > gpurun_out/topo_n$N.txt
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py --match 20000 --match-big 100000 > gpurun_out/multi_check_n$N.json 2> gpurun_out/multi_check_n$N.err; echo "multi_check rc=$?"; tail -5 gpurun_out/multi_check_n$N.err; cat gpurun_out/multi_check_n$N.json
timeout -s KILL 600 python scripts/slab_single_process.py $N 512 > gpurun_out/slab_sp_n$N.json 2> gpurun_out/slab_sp_n$N.err; echo "single-process rc=$?"; tail -5 gpurun_out/slab_sp_n$N.err; cat gpurun_out/slab_sp_n$N.json
