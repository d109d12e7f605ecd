mkdir -p gpurun_out
python scripts/profile_static.py > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_static.csv python scripts/profile_static.py > gpurun_out/ncu_static.log 2>&1; echo "ncu launch list rc=$?"
# 4 frames x 21 GEMM launches: capture the last frame's
ncu --set full --clock-control none --import-source on -k regex:"spconv_t[sc]_kernel" -s 63 -c 21 -f -o gpurun_out/prof_gemm_static \
    python scripts/profile_static.py > gpurun_out/ncu_full_static.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/
