#!/bin/bash
# round 2, visit B (N GPUs): new sharded matcher tests, NCCL slab/matcher checks, the new bench line at N
N=${1:-2}
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_match.py tests/test_gpu_slab.py -m gpu -x -q > gpurun_out/pytest_b.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_b.log
if [ "$N" -gt 1 ]; then
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_n$N.err
else
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_n$N.err
fi
grep "^{" gpurun_out/bench_n$N.json > gpurun_out/bench_n$N.clean.json; python scripts/show_bench.py gpurun_out/bench_n$N.clean.json | head -3
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n$N.clean.json"))
print(json.dumps(d.get("extra"), indent=1)[:3000])
print({k:v for k,v in (d.get("match") or {}).items() if k in ("ms","pairs_per_s","sampled_exact_equal","cpu_baseline")})
print(d["e2e"]["latency_ms_single_volume"], d.get("issue_roofline"), d.get("cpu_baseline"))
PY
