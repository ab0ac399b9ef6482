#!/bin/bash
# Build an experimental variant of the library next to the product one:
#   tools/build_variant.sh NAME -DWICP_CAP=184 -DFITW_MIN_BLOCKS=5   ->  fusion4landslide_b200/libf4l_b200_NAME.so
# (select it with F4L_LIB=... ; A/B experiments only, the product build is fusion4landslide_b200/build.py)
set -e
name=$1; shift
here=$(cd "$(dirname "$0")/.." && pwd)
out=$here/fusion4landslide_b200/build_$name
mkdir -p $out
for f in $here/fusion4landslide_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
    --expt-relaxed-constexpr "$@" -c $f -o $out/$(basename ${f%.cu}).o &
done
wait
nvcc -shared -Wno-deprecated-gpu-targets -o $here/fusion4landslide_b200/libf4l_b200_$name.so $out/*.o -lcudart
echo built libf4l_b200_$name.so
