#!/bin/bash
# ncu --set full of the last frame's gather-GEMM launches (static plan, config A): tag = $1
tag=${1:-r2}
mkdir -p gpurun_out
python scripts/profile_static.py > gpurun_out/${tag}_plain.log 2>&1 || { tail -5 gpurun_out/${tag}_plain.log; exit 1; }
# 4 frames x 21 GEMM launches: skip 3 frames
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"spconv_t[sc]_kernel|spconv_rg_kernel" -s 63 -c 21 \
    -f -o gpurun_out/prof_${tag}_gemm python scripts/profile_static.py > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/prof_${tag}_gemm.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_gemm_raw.csv 2>/dev/null
python scripts/ncu_extract.py gpurun_out/prof_${tag}_gemm_raw.csv gpurun_out/${tag}_ncu_full_spconv.csv gpurun_out/${tag}_traffic.json "gather-GEMM kernels x21 (one frame, config A, static plan)"
