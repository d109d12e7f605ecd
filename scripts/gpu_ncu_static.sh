mkdir -p gpurun_out
python scripts/profile_static.py > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_static.csv python scripts/profile_static.py > gpurun_out/ncu_static.log 2>&1; echo "ncu rc=$?"
