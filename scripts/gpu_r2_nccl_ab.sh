#!/bin/bash
# NCCL transport settings A/B for the slab leg at N GPUs
N=${1:-8}
mkdir -p gpurun_out; : > gpurun_out/slab_ab_n$N.txt
# Every variant runs in its OWN process group and is killed as a group after 120 s: round 2 lost its remaining GPU allowance to
# a variant that hung inside NCCL while `timeout` only killed the torchrun launcher and the ranks kept the output pipe open.
run() { # tag, env...
  tag=$1; shift
  out=$(mktemp)
  env "$@" setsid python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/slab_bench.py --tag "$tag" > "$out" 2>/dev/null &
  pid=$!
  for i in $(seq 1 120); do kill -0 $pid 2>/dev/null || break; sleep 1; done
  if kill -0 $pid 2>/dev/null; then kill -KILL -- -$pid 2>/dev/null; echo "{\"tag\": \"$tag\", \"hung\": true}" | tee -a gpurun_out/slab_ab_n$N.txt; sleep 5; fi
  grep '^{' "$out" | tee -a gpurun_out/slab_ab_n$N.txt; rm -f "$out"
}
run default X=1
run p2p16 NCCL_MIN_P2P_NCHANNELS=16
run p2p32 NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32
# run cudamemcpy NCCL_P2P_USE_CUDA_MEMCPY=1    # hangs with NCCL 2.28.9 on this box (round 2)
run group1 S3D_SLAB_GROUP=1
run group6 S3D_SLAB_GROUP=6
run group6_p2p32 S3D_SLAB_GROUP=6 NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32
