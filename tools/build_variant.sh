#!/bin/bash
# experiment builds: tools/build_variant.sh NAME -DFLAG... -> _variants/libauncel_NAME.so (use with AUNCEL_LIB=...)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p _variants/obj_$name
for f in coarse scan tcfilter tcfilter2 tcfilter3 merge merge_tables kmeans range shards shard_rounds index c_api; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false \
    -Xcompiler -fPIC,-O2,-fno-fast-math -ccbin /usr/bin/g++ "$@" -c auncel_b200/csrc/$f.cu -o _variants/obj_$name/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -shared -o _variants/libauncel_$name.so _variants/obj_$name/*.o -ccbin /usr/bin/g++
rm -rf _variants/obj_$name
echo _variants/libauncel_$name.so
