#!/bin/bash
# Launch list + full ncu capture of the candidate kernels at N x N (default 100k), paths 3 and 4.
mkdir -p gpurun_out
N=${1:-100000}
for path in 3 4; do
  timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/match_launches_p$path.csv python scripts/profile_match.py $N $path > gpurun_out/match_list_p$path.log 2>&1; echo "list p$path rc=$?"
  timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:topk_kernel -c 1 -f -o gpurun_out/full_match_p$path python scripts/profile_match.py $N $path > gpurun_out/ncu_match_p$path.log 2>&1; echo "ncu p$path rc=$?"
done
