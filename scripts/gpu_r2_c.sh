#!/bin/bash
# round 2, visit C: GPU suite after the sync-free sparse stage / persistent descriptor CTAs, bench at N=1 (short CPU legs)
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q -k "not 512 and not 10k" > gpurun_out/pytest_c.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_c.log
timeout 900 python bench.py --steps 20 --cpu-sample 128 --cpu-match-n 2000 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_c.err
python scripts/show_bench.py gpurun_out/bench_c.json | head -24
python - <<PY
import json
d=json.load(open("gpurun_out/bench_c.json"))
print("e2e", d["e2e"]["ms_per_step"], d["e2e"]["latency_ms_single_volume"], d["e2e"]["host_wall_ms_per_step"][:12])
s=d["extra"]["slab"]; print("slab", s["ms_per_volume"], s["latency_ms_single_volume"], s["e2e"]["ms_per_volume"], s["phases_ms_rank0"])
print("match_sharded", d["extra"]["match_sharded"]["ms"], d["match"]["ms"], d["match"].get("cpu_baseline"))
PY
S3D_E2E_THREADS=3 timeout 600 python bench.py --steps 20 --no-cpu-baseline --match-n 0 --no-extra > gpurun_out/bench_c_thr3.json 2>/dev/null; python - <<PY
import json
d=json.load(open("gpurun_out/bench_c_thr3.json")); print("threads=3 e2e", d["e2e"]["ms_per_step"], "value", d["ms_per_step"])
PY
