#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'describe_kernel' -c 1 -f -o /tmp/ncu/full_desc python scripts/profile_step.py 512 1 > gpurun_out/ncu_desc.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py /tmp/ncu/full_desc.ncu-rep --src 200 > gpurun_out/ncu_desc_summary.txt 2>&1
ncu -i /tmp/ncu/full_desc.ncu-rep --page source --csv --print-source sass > /tmp/ncu/desc_sass.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open('/tmp/ncu/desc_sass.csv')))
# keep: address/source, instructions executed, samples
hdr = None
out = []
for r in rows:
    if hdr is None:
        if 'Source' in r and any('Instructions Executed' in c for c in r):
            hdr = r
            si = r.index('Source'); ii = [i for i, c in enumerate(r) if c == 'Instructions Executed'][0]
            ti = [i for i, c in enumerate(r) if c.startswith('Thread Instructions Executed')]
            wi = [i for i, c in enumerate(r) if 'Warp Stall Sampling (All' in c]
        continue
    try:
        out.append((r[si], r[ii], r[ti[0]] if ti else '', r[wi[0]] if wi else ''))
    except Exception:
        pass
with open('gpurun_out/desc_sass.txt', 'w') as f:
    for o in out:
        f.write('\t'.join(o) + '\n')
print(len(out), 'sass lines')
PY
du -sh gpurun_out
