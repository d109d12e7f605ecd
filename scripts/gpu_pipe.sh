mkdir -p gpurun_out
export BEVFRONT_LIB=$PWD/bevfusion_3d_object_detection_b200/lib/libbevfront_b200_pipe.so
timeout 600 python -m pytest tests/test_spconv_gpu.py -m gpu -q -k "ts_kernel or bf16 or encoder or epilogue" -p no:cacheprovider --timeout 120 > gpurun_out/t_pipe.log 2>&1; rc=$?
echo "pipelined tests rc=$rc"; tail -5 gpurun_out/t_pipe.log | cut -c1-300
timeout 200 python scripts/profile_layers.py bf16 1 2>&1 | grep -v "^\[" | grep "64->  64\|128-> 128\|64-> 128\|total"
unset BEVFRONT_LIB
timeout 200 python scripts/profile_layers.py bf16 1 2>&1 | grep -v "^\[" | grep "64->  64\|128-> 128\|64-> 128\|total"
