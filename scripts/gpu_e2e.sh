#!/bin/bash
mkdir -p gpurun_out
for v in "S3D_E2E_ASYNC_D2H=0" "S3D_E2E_ASYNC_D2H=1"; do
  env $v timeout 600 python bench.py --steps 10 --warmup 3 --match-n 0 --no-cpu-baseline 2> gpurun_out/bench_e2e.err > gpurun_out/bench_e2e_$v.json
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_e2e_$v.json"))
print("$v", "value", round(d["value"],1), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],2), d["e2e"]["host_wall_ms_per_step"], d["e2e"]["last_step_split_ms"])
PY
done
