#!/bin/bash
# ncu --set full of one register-gather launch (16 -> 16, module path, 4th encoder pass)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spconv_rg_kernel" -s 19 -c 2 \
    -f -o gpurun_out/prof_rg python scripts/profile_gemm.py > gpurun_out/ncu_rg.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_rg.log
ncu -i gpurun_out/prof_rg.ncu-rep --page raw --csv > gpurun_out/prof_rg_raw.csv 2>/dev/null
ls -la gpurun_out/prof_rg*
