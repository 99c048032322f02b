#!/bin/bash
# Builds an experiment variant of the library with extra -D flags on kreg_kbuild.cu only (Nat = 9 instantiation):
#   bash tools/build_variant.sh NAME -DKREG_EMIT_NJB=4 ...   ->  tools/bin/lib_NAME.so
set -e
name=$1; shift
cd "$(dirname "$0")/../mlatom_b200/csrc"
mkdir -p build/var ../../tools/bin
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ \
  -DKREG_ONLY_NAT9 "$@" -c kreg_kbuild.cu -o build/var/kbuild_$name.o
objs=$(ls build/*.o | grep -v kreg_kbuild.o)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/bin/lib_$name.so $objs build/var/kbuild_$name.o -lcusolver -lcublas -ccbin /usr/bin/g++
echo built tools/bin/lib_$name.so
