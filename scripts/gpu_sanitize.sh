#!/bin/bash
# compute-sanitizer over small extractions (memcheck: out-of-bounds / misaligned; racecheck: shared-memory hazards in the
# cp.async rings, the descriptor queues and the orientation double buffer).  Only the summaries travel back.
mkdir -p gpurun_out; rm -f gpurun_out/sanitize.txt
for tool in memcheck racecheck; do
  for n in 64 48; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 5 python scripts/profile_step.py $n 1 > /tmp/san_${tool}_$n.log 2>&1
    echo "== $tool $n rc=$?" >> gpurun_out/sanitize.txt
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|keypoints" /tmp/san_${tool}_$n.log | head -12 >> gpurun_out/sanitize.txt
  done
done
cat gpurun_out/sanitize.txt
