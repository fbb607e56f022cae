#!/bin/bash
# One GPU-box visit at the end of a work session: every GPU parity test, smoke, both bench arms, the ncu launch list of
# one 512^3 step, DRAM traffic of the blur launches, and full ncu captures of the top kernels.  Only text/JSON summaries are
# written under gpurun_out/ (the .ncu-rep files stay in /tmp on the box: gpurun_out is limited to 64 MiB).
mkdir -p gpurun_out /tmp/ncu
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout -s KILL 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python scripts/show_bench.py gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py 512 1 > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"blur_march|blur_xy|blur_x_kernel" -c 200 -f -o /tmp/ncu/traffic_blur python scripts/profile_step.py 512 1 > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
python scripts/make_traffic.py /tmp/ncu/traffic_blur.ncu-rep gpurun_out/traffic.json > /dev/null 2>&1; echo "traffic rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'blur_march|blur_xy' -c 12 -f -o /tmp/ncu/full_blur python scripts/profile_step.py 512 1 > gpurun_out/ncu_blur.log 2>&1; echo "ncu blur rc=$?"
python scripts/ncu_summary.py /tmp/ncu/full_blur.ncu-rep --src 12 > gpurun_out/ncu_blur_summary.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'describe_kernel|orient_kernel|orient_exact_kernel' -c 3 -f -o /tmp/ncu/full_sparse python scripts/profile_step.py 512 1 > gpurun_out/ncu_sparse.log 2>&1; echo "ncu sparse rc=$?"
python scripts/ncu_summary.py /tmp/ncu/full_sparse.ncu-rep --src 40 > gpurun_out/ncu_sparse_summary.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'detect_kernel|scan_kernel' -c 4 -f -o /tmp/ncu/full_detect python scripts/profile_step.py 512 1 > gpurun_out/ncu_detect.log 2>&1; echo "ncu detect rc=$?"
python scripts/ncu_summary.py /tmp/ncu/full_detect.ncu-rep --src 20 > gpurun_out/ncu_detect_summary.txt 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'maxabs_kernel|normalize_kernel|downsample_kernel' -c 3 -f -o /tmp/ncu/full_small python scripts/profile_step.py 512 1 > gpurun_out/ncu_small.log 2>&1; echo "ncu small rc=$?"
python scripts/ncu_summary.py /tmp/ncu/full_small.ncu-rep > gpurun_out/ncu_small_summary.txt 2>&1
bash scripts/gpu_sanitize.sh > /dev/null 2>&1; echo "sanitize rc=$?"; grep -E "^==|SUMMARY" gpurun_out/sanitize.txt
du -sh gpurun_out
