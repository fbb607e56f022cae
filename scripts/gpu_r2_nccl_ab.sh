#!/bin/bash
# NCCL transport settings A/B for the slab leg at N GPUs
N=${1:-8}
mkdir -p gpurun_out; : > gpurun_out/slab_ab_n$N.txt
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/slab_bench.py --tag "$tag" 2>/dev/null | grep '^{' | tee -a gpurun_out/slab_ab_n$N.txt
}
run default X=1
run p2p16 NCCL_MIN_P2P_NCHANNELS=16
run p2p32 NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32
run cudamemcpy NCCL_P2P_USE_CUDA_MEMCPY=1
run group1 S3D_SLAB_GROUP=1
run group6 S3D_SLAB_GROUP=6
run group6_p2p32 S3D_SLAB_GROUP=6 NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32
