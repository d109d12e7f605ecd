mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_static_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/t_some.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/t_some.log
