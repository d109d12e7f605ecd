mkdir -p gpurun_out
python scripts/profile_static.py > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:"spconv_tc_kernel" -s 42 -c 21 -o gpurun_out/prof_tc_final python scripts/profile_static.py > gpurun_out/ncu_tc_final.log 2>&1; echo "ncu rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --mode eager > gpurun_out/plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --mode eager > gpurun_out/ncu_launch_final.log 2>&1; echo "ncu launches rc=$?"
