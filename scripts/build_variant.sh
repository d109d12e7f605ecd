#!/bin/bash
# build a variant of the library next to the default one: usage build_variant.sh <tag> "<nvcc defines>" [files...]
# -> bevfusion_3d_object_detection_b200/lib/libbevfront_b200_<tag>.so (select it with BEVFRONT_LIB=...)
set -e
tag="$1"; defs="$2"; shift 2
files="${@:-spconv_tc.cu}"
P=bevfusion_3d_object_detection_b200
mkdir -p $P/build/var_$tag
objs=""
for f in $P/csrc/*.cu; do
  b=$(basename $f .cu)
  if echo " $files " | grep -q " $b.cu "; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -Iinclude -I$P/csrc $defs -c $f -o $P/build/var_$tag/$b.o &
    objs="$objs $P/build/var_$tag/$b.o"
  else
    objs="$objs $P/build/$b.o"
  fi
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $P/lib/libbevfront_b200_$tag.so $objs
echo built $P/lib/libbevfront_b200_$tag.so
