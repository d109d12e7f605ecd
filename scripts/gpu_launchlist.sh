#!/bin/bash
# ncu launch list of the static plan (last frame), tag = $1
tag=${1:-r2}
mkdir -p gpurun_out
python scripts/profile_static.py > gpurun_out/${tag}_plain.log 2>&1 || { tail -5 gpurun_out/${tag}_plain.log; exit 1; }
n=$(grep launches_per_frame gpurun_out/${tag}_plain.log | awk '{print $2}')
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${tag}_launches.csv \
    python scripts/profile_static.py > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$? launches_per_frame=$n"
python scripts/launch_table.py gpurun_out/${tag}_launches.csv $n
