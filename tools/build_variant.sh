#!/bin/bash
# Build a variant of the library with extra nvcc defines for the tensor-core ensemble kernels (A/B experiments, see
# tools/ab_bench.sh):   tools/build_variant.sh polyAA -DNPHM_POLY_MASK=0xAA   ->  nphm_b200/libnphm_b200_polyAA.so
set -e
NAME=$1; shift
cd "$(dirname "$0")/../nphm_b200/csrc"
make -s all > /dev/null
for f in tc_ensemble tc_ensemble_v8; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC "$@" -c $f.cu -o build/${f}__$NAME.o
done
OBJS=$(ls build/*.o | grep -v "tc_ensemble")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libnphm_b200_$NAME.so $OBJS build/tc_ensemble__$NAME.o build/tc_ensemble_v8__$NAME.o -lcudart
rm -f build/tc_ensemble__$NAME.o build/tc_ensemble_v8__$NAME.o
echo built nphm_b200/libnphm_b200_$NAME.so
