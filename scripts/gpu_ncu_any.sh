#!/bin/bash
# ncu --set full of kernels matching $1 while running "$5..." : $2 = tag, $3 = skip, $4 = count
re="$1"; tag="$2"; skip="$3"; cnt="$4"; shift 4
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$re" -s $skip -c $cnt -f -o gpurun_out/prof_$tag "$@" > gpurun_out/ncu_$tag.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
