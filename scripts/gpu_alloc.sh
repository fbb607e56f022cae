#!/bin/bash
# block cache against the stream-ordered pool: parity tests, A/B timing, the e2e variants that overlap handle lifetimes
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_sparse.py tests/test_gpu_slab.py tests/test_facade.py -m gpu -q -x > gpurun_out/pytest_alloc.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_alloc.log
python scripts/diag_e2e_host.py 2>&1 | tail -1
S3D_ALLOC=pool python scripts/diag_e2e_host.py 2>&1 | tail -1
for v in "S3D_ALLOC=pool S3D_E2E_ASYNC_D2H=0" "S3D_ALLOC=cache S3D_E2E_ASYNC_D2H=0" "S3D_ALLOC=cache S3D_E2E_ASYNC_D2H=1" "S3D_ALLOC=cache S3D_E2E_ASYNC_D2H=1"; do
  env $v timeout 600 python bench.py --steps 20 --warmup 3 --match-n 0 --no-cpu-baseline 2> gpurun_out/bench_alloc.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$v', 'value', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), d['e2e']['host_wall_ms_per_step'])"
done
