#!/bin/bash
# usage: tools/build_variant.sh <name> [extra nvcc -D flags...]   -> nvdiffrecmc_b200/lib/variants/<name>.so
set -e
cd "$(dirname "$0")/../nvdiffrecmc_b200/csrc"
name=$1; shift
mkdir -p ../lib/variants/obj_$name
for f in core elementwise denoise bvh envshade lossmesh light raster; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC "$@" -c $f.cu -o ../lib/variants/obj_$name/$f.o &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../lib/variants/$name.so ../lib/variants/obj_$name/*.o
rm -rf ../lib/variants/obj_$name
echo built $name
