#!/bin/bash
# bench lines at several N on one multi-GPU box:  gpu_r2_scale.sh "4 8"
mkdir -p gpurun_out
for N in $1; do
  timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"; tail -3 gpurun_out/bench_n$N.err
  grep "^{" gpurun_out/bench_n$N.json > gpurun_out/bench_n$N.clean.json
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_n$N.clean.json"))
print("N=$N value",round(d["value"]),round(d["ms_per_step"],2),"e2e",round(d["e2e"]["value"]),round(d["e2e"]["ms_per_step"],2),"lat",round(d["e2e"]["latency_ms_single_volume"],1))
s=d["extra"]["slab"]; print("  slab value",round(s["value"]),round(s["ms_per_volume"],2),"e2e",round(s["e2e"]["value"]),round(s["e2e"]["ms_per_volume"],2), "eq", s["equal_to_single_gpu_keypoints"])
for r in s["phases_per_rank"]: print("   ", r)
m=d["extra"]["match_sharded"]; print("  match_sharded ms",round(m["ms"],1),"eq",m["equal_to_single_gpu_outputs"],"bytes",m["nccl_bytes_sent_per_match_rank0"], m["clocks"])
print("  match ms", round(d["match"]["ms"],1), d["match"]["clocks"])
print("  per_rank", d["per_rank"])
PY
done
