#!/bin/bash
# Build an instrumented variant of the library into probes/_libs/libpylom_<name>.so (select it with PL_LIBPATH).
# usage: probes/build_variant.sh <name> <extra nvcc flags...>
set -e
name=$1; shift
here=$(cd "$(dirname "$0")" && pwd)
src=$here/../pyloworder_b200/csrc
out=$here/_libs; mkdir -p $out/obj_$name
for f in api center gemm gemm_tn caqr tsqr_small svd_small comm; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O3 "$@" -I$src -c $src/$f.cu -o $out/obj_$name/$f.o &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $out/libpylom_$name.so $out/obj_$name/*.o -lcudart -ldl
rm -rf $out/obj_$name
echo built $out/libpylom_$name.so
