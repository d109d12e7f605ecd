#!/bin/bash
# ncu --set full of selected kernels of the static plan: $1 = kernel-name regex, $2 = tag, $3 = skip, $4 = count
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${3:-0} -c ${4:-4} \
    -f -o gpurun_out/prof_$2 python scripts/profile_static.py > gpurun_out/ncu_$2.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/prof_$2.ncu-rep --page raw --csv > gpurun_out/prof_$2_raw.csv 2>/dev/null
ls -la gpurun_out/prof_$2*
