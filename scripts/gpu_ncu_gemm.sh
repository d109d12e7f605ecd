mkdir -p gpurun_out
# TS launches per encoder pass: 5 x 16ch, 4 x 32, 4 x 64, 4 x 128; skip 3 passes + the 16ch layers + 3 of the 32ch
timeout 600 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section LaunchStats \
    --clock-control none --import-source on -k regex:"spconv_ts_kernel" -s 59 -c 3 \
    -f -o gpurun_out/prof_gemm2 python scripts/profile_gemm.py > gpurun_out/ncu_gemm.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_gemm.log
ls -la gpurun_out/*.ncu-rep
