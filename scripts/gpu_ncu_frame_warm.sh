#!/bin/bash
# like gpu_ncu_frame.sh but WITHOUT ncu's cache flush before every kernel (--cache-control none): the DRAM traffic the
# frame's GEMM launches see when they run back to back, as they do in the frame (each layer's input was just written by the
# previous one, the level's rulebook was read by the previous layer); tag = $1
tag=${1:-r2}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --cache-control none -k regex:"spconv_t[sc]_kernel|spconv_rg_kernel" -s 63 -c 21 \
    -f -o gpurun_out/prof_${tag}_gemm_warm python scripts/profile_static.py > gpurun_out/${tag}_ncu_full_warm.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/prof_${tag}_gemm_warm.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_gemm_warm_raw.csv 2>/dev/null
python scripts/ncu_extract.py gpurun_out/prof_${tag}_gemm_warm_raw.csv gpurun_out/${tag}_ncu_full_spconv_warm.csv gpurun_out/${tag}_traffic_warm.json "gather-GEMM kernels x21 (one frame, config A, static plan; no cache flush between kernels)"
