mkdir -p gpurun_out
python scripts/profile_small.py > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"bev_pool_runs|rank_scan_emit|strided_mark|subm_rulebook|strided_rulebook|vox_|to_bev|transpose" -s 60 -c 40 \
    -o gpurun_out/prof_small python scripts/profile_small.py > gpurun_out/ncu_small.log 2>&1; echo "ncu rc=$?"
