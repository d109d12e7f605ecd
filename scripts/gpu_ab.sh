#!/bin/bash
# A/B of prebuilt library variants (scripts/build_variant.sh): usage gpu_ab.sh <tag|default> ...   [TESTS=1 runs the bf16 GEMM tests first]
mkdir -p gpurun_out
P=$PWD/bevfusion_3d_object_detection_b200/lib
if [ -n "$TESTS" ]; then
  timeout 900 python -m pytest tests/test_spconv_gpu.py tests/test_static_gpu.py -m gpu -q -x -p no:cacheprovider --timeout 300 > gpurun_out/t_ab.log 2>&1
  echo "tests rc=$?"; tail -3 gpurun_out/t_ab.log | cut -c1-300
fi
for rep in 1 2; do
for tag in "$@"; do
  if [ "$tag" = default ]; then lib=$P/libbevfront_b200.so; else lib=$P/libbevfront_b200_$tag.so; fi
  BEVFRONT_LIB=$lib TAG="$tag" timeout 300 python scripts/layer_times.py 2>&1 | grep -E "layer_us|gemm_ms"
done
done
