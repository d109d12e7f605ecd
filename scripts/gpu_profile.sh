mkdir -p gpurun_out
set -x
python scripts/profile_layers.py bf16 > gpurun_out/layers_bf16.log 2>&1; echo "layers rc=$?"
timeout 300 python -m pytest tests/test_spconv_gpu.py -m gpu -q -k "training" -p no:cacheprovider > gpurun_out/t_train.log 2>&1; echo "train rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"spconv_tc_kernel|bev_pool_fused_fwd|vox_insert|vox_rank|vox_gather_mean" -s 300 -c 40 \
    -o gpurun_out/prof_r1a python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
