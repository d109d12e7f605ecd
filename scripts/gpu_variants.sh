#!/bin/bash
# build the library with extra nvcc defines and time the GEMM layers: usage gpu_variants.sh "<tag>=<defs>" ...
mkdir -p gpurun_out
for spec in "$@"; do
  tag="${spec%%=*}"; defs="${spec#*=}"
  BEVFRONT_NVCC_EXTRA="$defs" python -m bevfusion_3d_object_detection_b200.build -f > /dev/null 2>&1 || { echo "build failed $tag"; continue; }
  TAG="$tag" timeout 300 python scripts/layer_times.py 2>&1 | grep -E "layer_us|gemm_ms" 
done
python -m bevfusion_3d_object_detection_b200.build -f > /dev/null 2>&1
