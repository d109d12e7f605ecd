mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_spconv_gpu.py -m gpu -q -k "ts_kernel or bf16 or encoder or epilogue" -p no:cacheprovider --timeout 120 > gpurun_out/t_ts.log 2>&1; rc=$?
echo "ts tests rc=$rc"; tail -40 gpurun_out/t_ts.log | cut -c1-300
if [ $rc -ne 0 ]; then exit 1; fi
for v in ${VARIANTS:-0 1}; do timeout 200 python scripts/profile_layers.py bf16 $v 2>&1 | grep -v "^\[" | tail -24; done
if [ -f bevfusion_3d_object_detection_b200/lib/libbevfront_b200_prof.so ]; then
  BEVFRONT_LIB=$PWD/bevfusion_3d_object_detection_b200/lib/libbevfront_b200_prof.so timeout 300 python scripts/prof_ts.py 2>&1 | tail -16
fi
