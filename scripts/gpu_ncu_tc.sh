mkdir -p gpurun_out
set -x
python scripts/profile_layers.py bf16 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"spconv_tc_kernel" -s 63 -c 21 \
    -o gpurun_out/prof_tc python scripts/profile_layers.py bf16 > gpurun_out/ncu_tc.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out
