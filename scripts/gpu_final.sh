mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider --timeout 300 > gpurun_out/t_all.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/t_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-700
python scripts/profile_static.py > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_static.csv python scripts/profile_static.py > gpurun_out/ncu_static.log 2>&1; echo "ncu launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"spconv_t[sc]_kernel" -s 63 -c 21 -f -o gpurun_out/prof_gemm_static \
    python scripts/profile_static.py > gpurun_out/ncu_full_static.log 2>&1; echo "ncu full rc=$?"
