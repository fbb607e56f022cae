#!/bin/bash
# round 2, visit D (2 GPUs): tests touched by grouping / facade / describe grid; 2-GPU bench incl. group-size A/B
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests/test_gpu_slab.py tests/test_facade.py tests/test_gpu_sparse.py tests/test_gpu_match.py -m gpu -x -q > gpurun_out/pytest_d.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_d.log
for G in 1 3 6; do
S3D_SLAB_GROUP=$G timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$G bench.py --gpus 2 --steps 10 --warmup 3 --match-n 0 > gpurun_out/bench_n2_g$G.json 2> gpurun_out/bench_n2_g$G.err; echo "bench group=$G rc=$?"; tail -2 gpurun_out/bench_n2_g$G.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n2_g$G.json"))
s=d["extra"]["slab"]; print("group=$G value ms",round(d["ms_per_step"],2),"e2e",round(d["e2e"]["ms_per_step"],2),"| slab ms",round(s["ms_per_volume"],2),"lat",round(s["latency_ms_single_volume"],2),"e2e",round(s["e2e"]["ms_per_volume"],2), s["phases_per_rank"])
PY
done
