mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_static_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/t_static.log 2>&1; echo "static rc=$?"
tail -30 gpurun_out/t_static.log
