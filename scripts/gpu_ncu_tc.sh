mkdir -p gpurun_out
python scripts/profile_static.py > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"spconv_tc_kernel" -s 42 -c 21 -o gpurun_out/prof_tc_final2 python scripts/profile_static.py > gpurun_out/ncu_tc_final2.log 2>&1; echo "ncu rc=$?"
