#!/bin/bash
# Full ncu captures of octave 0's blur launches and the sparse kernels; only the text summaries travel back.
mkdir -p gpurun_out /tmp/ncu
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'blur_march_kernel|blur_xyc_kernel|blur_xy_kernel' -c 12 -f -o /tmp/ncu/full_blur python scripts/profile_step.py 512 1 > gpurun_out/ncu_blur.log 2>&1; echo "ncu blur rc=$?"
python scripts/ncu_summary.py /tmp/ncu/full_blur.ncu-rep --src 14 > gpurun_out/ncu_blur_summary.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'describe_kernel|orient_kernel|orient_exact_kernel' -c 3 -f -o /tmp/ncu/full_sparse python scripts/profile_step.py 512 1 > gpurun_out/ncu_sparse.log 2>&1; echo "ncu sparse rc=$?"
python scripts/ncu_summary.py /tmp/ncu/full_sparse.ncu-rep --src 40 > gpurun_out/ncu_sparse_summary.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'detect_kernel|scan_kernel' -c 4 -f -o /tmp/ncu/full_detect python scripts/profile_step.py 512 1 > gpurun_out/ncu_detect.log 2>&1; echo "ncu detect rc=$?"
python scripts/ncu_summary.py /tmp/ncu/full_detect.ncu-rep --src 25 > gpurun_out/ncu_detect_summary.txt 2>&1
# extra raw metrics of the Z pass (memory system detail)
ncu -i /tmp/ncu/full_blur.ncu-rep --page raw --csv > /tmp/ncu/blur_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open('/tmp/ncu/blur_raw.csv')))
hdr = rows[0]
keep = [i for i, h in enumerate(hdr) if any(k in h for k in ("Kernel Name", "dram__", "lts__t_sector", "lts__t_bytes", "l1tex__t_bytes", "lts__average", "dram__cycles", "hit_rate", "gpu__time_duration", "fbpa", "lts__d_sectors"))]
with open('gpurun_out/ncu_blur_mem.csv', 'w') as f:
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[i] for i in keep])
PY
du -sh gpurun_out
