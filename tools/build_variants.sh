#!/bin/bash
# Build tuning variants of libhslam_b200.so into build/variants/ (git-ignored, shipped by gpurun).
# usage: tools/build_variants.sh name1:"-DFLAG=.. -DFLAG2=.." name2:"..." ...
set -e
cd "$(dirname "$0")/../h-slam_b200/csrc"
mkdir -p ../../build/variants
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  $NVCC $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $flags -Xptxas -v -c ba.cu -o ../../build/variants/ba_$name.o 2> ../../build/variants/ba_$name.ptxas.log
  $NVCC $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $flags -c tracker.cu -o ../../build/variants/tracker_$name.o 2>/dev/null
  $NVCC $ARCH -shared -o ../../build/variants/libhslam_b200_$name.so context.o pyramid.o ../../build/variants/tracker_$name.o ../../build/variants/ba_$name.o -lcudart_static -ldl -lpthread -lrt
  echo "built $name ($flags): $(grep -A2 'ba_linearize_kernelILb1' ../../build/variants/ba_$name.ptxas.log | grep -E 'Used' | head -1)"
done
