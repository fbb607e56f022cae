#!/bin/bash
# Full ncu capture of one kernel of a 512^3 extraction:  gpu_ncu_kernel.sh <kernel regex> [launch-skip] [count]
mkdir -p gpurun_out
K=$1; S=${2:-0}; C=${3:-1}
N=$(echo -n "$K" | tr -c 'A-Za-z0-9_' '_')
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c $C -f -o gpurun_out/full_$N python scripts/profile_step.py ${SIZE:-512} 1 > gpurun_out/ncu_$N.log 2>&1; echo "ncu $K rc=$?"
